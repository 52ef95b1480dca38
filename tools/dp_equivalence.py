'''
Data-parallel equivalence on real GPUs: N-rank loss and
gradients == 1-rank on the concatenated batch, and replicas stay identical.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dp_equivalence.py [--out gpurun_out/dp_equivalence.json]

Checks (dropout off, identical weights):
  1. mean loss over ranks == 1-rank loss on the whole batch;
  2. every gradient tensor after the bucketed NCCL all-reduce, divided by the world size (the factor Adam
     applies), == the 1-rank gradient of the whole batch (relative L2; rows are independent, so the only
     differences are fp32 summation order and the bf16 rounding of per-rank partial weight gradients);
  3. three optimizer steps: weights equal to the 1-rank run's and bit-identical across ranks;
  4. ranks that start from DIFFERENT seeds (what happens without --seed) agree on rank 0's seed and, after
     `sync_replicas`, hold bit-identical variables after three more steps.
Rank 0 prints and writes one JSON object; exit code 1 on a failed check.
'''

import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist

    from composer_b200 import _lib, parallel
    from composer_b200.models.transformer import Transformer

    parser = argparse.ArgumentParser()
    parser.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'dp_equivalence.json'))
    parser.add_argument('--layers', type=int, default=3)
    parser.add_argument('--seq', type=int, default=384)
    parser.add_argument('--per-rank', type=int, default=2)
    args = parser.parse_args()

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    dist.init_process_group('nccl', device_id=device)

    vocab, E, heads, T, L = 390, 256, 16, args.seq, args.layers

    def make_model(seed):
        return Transformer(vocab, E, T, L, heads, False, 0.0, 0.02, 0.0, 0.0, 1e-5, True, True, device=device, seed=seed)

    rng = np.random.default_rng(2024)
    draw = rng.integers(0, vocab, size=(world * args.per_rank, T + 1))
    x_all, y_all = draw[:, :-1], draw[:, 1:]
    mine = slice(rank * args.per_rank, (rank + 1) * args.per_rank)
    checks = []

    def stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ---- 1, 2: loss and gradients ----
    dp_model = make_model(5)
    loss_sum, _ = dp_model.forward_loss(x_all[mine], y_all[mine], training=True)
    local_loss = loss_sum.clone() / (args.per_rank * T)
    got_world = dp_model.backward()                       # bucketed all-reduce overlapped with backward
    torch.cuda.synchronize()
    dist.all_reduce(local_loss)
    dp_loss = float(local_loss) / world
    dp_grads = {k: v / world for k, v in dp_model.get_gradients().items()}

    single = make_model(5)
    loss_sum, _ = single.forward_loss(x_all, y_all, training=True)
    _lib.call('cb200_zero_grads', single._engine, stream())
    _lib.call('cb200_backward', single._engine, -1, stream())     # no collective: the 1-rank reference
    torch.cuda.synchronize()
    ref_loss = float(loss_sum) / (world * args.per_rank * T)
    ref_grads = single.get_gradients()
    checks.append({'name': 'world size seen by backward()', 'got': got_world, 'ok': got_world == world})
    checks.append({'name': 'mean loss over ranks == 1-rank loss', 'got': dp_loss, 'ref': ref_loss,
                   'rel': abs(dp_loss - ref_loss) / abs(ref_loss), 'tol': 2e-6,
                   'ok': abs(dp_loss - ref_loss) <= 2e-6 * abs(ref_loss)})
    worst, worst_name = 0.0, ''
    for name, ref in ref_grads.items():
        rel = float(np.linalg.norm(dp_grads[name].astype(np.float64) - ref) / (np.linalg.norm(ref) + 1e-30))
        if rel > worst:
            worst, worst_name = rel, name
    checks.append({'name': 'all-reduced gradients / world == 1-rank gradients (worst tensor: %s)' % worst_name,
                   'rel': worst, 'tol': 2e-3, 'ok': worst <= 2e-3})

    # ---- 3: optimizer steps ----
    dp_model, single = make_model(5), make_model(5)
    for _ in range(3):
        dp_model.train_step(x_all[mine], y_all[mine], 1e-3)
        loss_sum, _ = single.forward_loss(x_all, y_all, training=True, step=single._adam_t)
        _lib.call('cb200_zero_grads', single._engine, stream())
        _lib.call('cb200_backward', single._engine, -1, stream())
        single.apply_gradients(1e-3, 1)
    torch.cuda.synchronize()
    a, b = dp_model._params, single._params
    # Adam's first steps move every coordinate by ~lr whatever the gradient's size, so a coordinate whose gradient is
    # numerically zero may move the other way: compare the mean update error in units of lr
    diff = float((a - b).abs().mean()) / 1e-3
    checks.append({'name': '3 Adam steps: mean |w_dp - w_1rank| / lr', 'got': diff, 'tol': 0.05, 'ok': diff <= 0.05})
    gathered = [torch.empty_like(a) for _ in range(world)]
    dist.all_gather(gathered, a)
    same = all(bool(torch.equal(gathered[0], g)) for g in gathered[1:])
    checks.append({'name': 'replicas bit-identical after 3 steps', 'ok': same})

    # ---- 4: different seeds per rank (no --seed) ----
    seed = parallel.agree_on_seed(1000 + 77 * rank)
    checks.append({'name': 'agree_on_seed adopts rank 0\'s seed', 'got': seed, 'ok': seed == 1000})
    stray = make_model(1000 + 77 * rank)                 # replicas built from different seeds ...
    stray.sync_replicas()                                 # ... are pulled onto rank 0's variables
    for _ in range(3):
        stray.train_step(x_all[mine], y_all[mine], 1e-3)
    torch.cuda.synchronize()
    gathered = [torch.empty_like(stray._params) for _ in range(world)]
    dist.all_gather(gathered, stray._params)
    same = all(bool(torch.equal(gathered[0], g)) for g in gathered[1:])
    checks.append({'name': 'replicas from different seeds bit-identical after sync_replicas + 3 steps', 'ok': same})

    ok = all(c['ok'] for c in checks)
    if rank == 0:
        report = {'tool': 'tools/dp_equivalence.py', 'world_size': world, 'layers': L, 'seq_len': T,
                  'per_rank_batch': args.per_rank, 'gpu': torch.cuda.get_device_name(device), 'ok': ok, 'checks': checks}
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, 'w') as handle:
            json.dump(report, handle, indent=1)
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()
    raise SystemExit(0 if ok else 1)


if __name__ == '__main__':
    main()
