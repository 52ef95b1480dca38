'''
Benchmark of the Composer Transformer hot path on B200 (contract: see the task
brief / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): Transformer training step, bf16 tensor
math, default_config.yml hyperparameters with window_size 2048, dropout 0.1,
32 sequences x 2048 event tokens per GPU, data parallel over N GPUs (weak
scaling, NCCL gradient all-reduce overlapped with backward), synthetic uniform
token ids, random-init weights.  A step = forward + loss + backward +
all-reduce + Adam over one batch.

Rank 0 prints ONE JSON line.  `value` is whole-job tokens/s with the batch
already resident in HBM; `e2e` is the same metric through the public
`Transformer.train_step` call with host (pinned) batches, host->device copies
and the loss read-back inside the timed region.  `--impl reference` times the
CPU restatement of the reference (oracle/; TensorFlow is not installable in
this image, so the reference's own code cannot run) on the host cores: same
`config`, dropout on, one sequence of the batch per step (`config_delta`).
Also on the line: `roofline` (attention backward, against the tensor peak and
the MUFU.EX2 rate), `generate` (configs[2], ids copied to the host inside the
timed region, its own `cpu_baseline`), `other_configs` (configs[0], [3], [4]).
'''

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'train tokens/s (Transformer training step bf16, seq len 2048, data-parallel)'
UNIT = 'tokens/s'
VOCAB = 390             # 128 + 128 + 32 velocity bins + 100 time shifts + 2 sustain (default dataset config)
SEQ_LEN = 2048
BATCH_PER_GPU = 32
MODEL = dict(embedding_size=256, decoder_layers_count=8, attention_head_count=16, window_size=SEQ_LEN,
             attention_dropout_rate=0.1, residual_dropout_rate=0.1)


def step_flops_per_token(layers, embedding, vocab, seq_len):
    '''Algorithmic FLOPs per trained token: 3 x forward, recompute not counted.'''

    forward = layers * 24 * embedding ** 2 + 2 * embedding * vocab + layers * 2 * embedding * (seq_len + 1)
    return 3 * forward


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as handle:
            peaks = json.load(handle)
        return peaks, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    '''Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs.'''

    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.device_index = device_index
        self.samples = []
        self.process = None
        self.thread = None

    def start(self):
        try:
            self.process = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.device_index), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.process = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.process.stdout:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) >= 8:
                self.samples.append(parts)

    def stop(self):
        if self.process is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.process.terminate()
        try:
            self.process.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.process.kill()
        clocks, reasons, sm_max, power = [], set(), None, 0.0
        for parts in self.samples:
            try:
                clocks.append(float(parts[1]))
                sm_max = float(parts[2])
                power = max(power, float(parts[3]))
            except ValueError:
                continue
            for name, value in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                   parts[4:8]):
                if value.lower().startswith('active'):
                    reasons.add(name)
        clocks.sort()
        # median over the upper half: the sampler also sees the idle moments around the timed region
        busy = clocks[len(clocks) // 2:] if clocks else []
        median = busy[len(busy) // 2] if busy else None
        return {'sm_mhz': median, 'sm_max_mhz': sm_max, 'reasons': sorted(reasons), 'samples': len(clocks),
                'power_w_max': power}


# ---------------------------------------------------------------------------
# CPU baseline: the oracle (a restatement of the reference's TensorFlow model)
# ---------------------------------------------------------------------------

def gpu_config(world, batch):
    '''The `config` object of both arms (the reference arm times a bounded sample of the same workload).'''

    return {'workload': 'BASELINE.json configs[1]: Transformer training step bf16, seq len 2048 '
                        '(default_config.yml hyperparameters, window_size 2048, dropout 0.1)',
            'global_batch': world * batch, 'per_gpu_batch': batch, 'seq_len': SEQ_LEN, 'vocab': VOCAB,
            'parallelism': 'dp%d' % world,
            'l2_policy': 'inputs larger than L2: each step streams >4 GB of activations'}


def oracle_dropout_masks(cfg, batch, seq_len, generator):
    '''Keep masks for every dropout site of the reference model (transformer.py:361, 444, 506, 794), drawn on the host.'''

    import torch

    def keep(shape, rate):
        return (torch.rand(shape, generator=generator) >= rate).to(torch.float32)

    heads, E = cfg.attention_head_count, cfg.embedding_size
    masks = {'embd': keep((batch, seq_len, E), cfg.residual_dropout_rate)}
    for layer in range(1, cfg.decoder_layers_count + 1):
        masks[(layer, 'attn_w')] = keep((batch, heads, seq_len, seq_len), cfg.attention_dropout_rate)
        masks[(layer, 'attn_resid')] = keep((batch, seq_len, E), cfg.residual_dropout_rate)
        masks[(layer, 'mlp')] = keep((batch, seq_len, E), cfg.residual_dropout_rate)
    return masks


def cpu_reference_step(batch, seq_len, steps, warmup, budget_seconds=120.0):
    '''
    Times the oracle's training step (forward with dropout 0.1 at every site, loss, autograd, Keras Adam) in fp32 on
    all host cores.  Mask generation is inside the timed region, as it is inside the reference's step.  At most
    `steps` timed steps, fewer when `budget_seconds` of CPU work are used up first (the mean is over what ran).
    '''

    import numpy as np
    import torch
    from oracle import transformer_oracle as oracle

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = oracle.OracleConfig(vocab_size=VOCAB, embedding_size=MODEL['embedding_size'], window_size=seq_len,
                              decoder_layers_count=MODEL['decoder_layers_count'],
                              attention_head_count=MODEL['attention_head_count'],
                              attention_dropout_rate=MODEL['attention_dropout_rate'],
                              residual_dropout_rate=MODEL['residual_dropout_rate'])
    weights = oracle.init_parameters(cfg, seed=0)
    params = oracle.to_torch(weights, torch.float32, requires_grad=True)
    optimizer_state = {name: (torch.zeros_like(p), torch.zeros_like(p)) for name, p in params.items()}
    rng = np.random.default_rng(1234)
    generator = torch.Generator().manual_seed(1234)
    times = []
    for step in range(warmup + steps):
        draw = rng.integers(0, VOCAB, size=(batch, seq_len + 1))
        x, y = draw[:, :-1], draw[:, 1:]
        start = time.perf_counter()
        masks = oracle_dropout_masks(cfg, batch, seq_len, generator)
        logits, _ = oracle.transformer_call(params, x, cfg, dropout_masks=masks)
        loss = oracle.sparse_categorical_crossentropy(y, logits)
        loss.backward()
        t = step + 1
        lr_t = 1e-3 * (1 - 0.999 ** t) ** 0.5 / (1 - 0.9 ** t)
        with torch.no_grad():
            for name, p in params.items():
                m, v = optimizer_state[name]
                m.mul_(0.9).add_(p.grad, alpha=0.1)
                v.mul_(0.999).addcmul_(p.grad, p.grad, value=0.001)
                p.sub_(lr_t * m / (v.sqrt() + 1e-7))
                p.grad = None
        elapsed = time.perf_counter() - start
        if step >= warmup:
            times.append(elapsed)
            if sum(times) > budget_seconds:
                break
    mean = sum(times) / len(times)
    return batch * seq_len / mean, threads, mean


def cpu_reference_generate(sequences, length):
    '''Times the oracle's KV-cache decode (transformer.py:735-770 `past=`) on the host cores: events/s.'''

    import numpy as np
    import torch
    from oracle import transformer_oracle as oracle

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = oracle.OracleConfig(vocab_size=VOCAB, embedding_size=MODEL['embedding_size'], window_size=1024,
                              decoder_layers_count=MODEL['decoder_layers_count'],
                              attention_head_count=MODEL['attention_head_count'],
                              attention_dropout_rate=0.0, residual_dropout_rate=0.0)
    weights = oracle.init_parameters(cfg, seed=0)
    rng = np.random.default_rng(99)
    prompt = rng.integers(0, VOCAB, size=(sequences, 1))
    uniforms = rng.random((sequences, length))
    start = time.perf_counter()
    oracle.generate(weights, prompt, length, cfg, temperature=1.0, dtype=torch.float32, uniforms=uniforms)
    seconds = time.perf_counter() - start
    return sequences * length / seconds, threads, seconds


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', '1'))
    batch = 1
    value, threads, seconds = cpu_reference_step(batch, SEQ_LEN, max(1, args.steps), max(0, min(args.warmup, 1)))
    sample = ('oracle/transformer_oracle.py (CPU restatement of composer/models/transformer.py, pinned to the '
              'reference\'s own model code under a TensorFlow stand-in; TensorFlow itself is not installable here) '
              'training step, fp32, dropout 0.1 at every site, 1 of the %d sequences x T=%d per step, up to %d timed steps '
              '(120 s budget) of %.1f s' % (args.batch, SEQ_LEN, max(1, args.steps), seconds))
    gen_value, _, gen_seconds = cpu_reference_generate(4, 256)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': seconds * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': gpu_config(world, args.batch),
        'config_delta': {'per_step_batch': batch, 'why': 'bounded sample: a 32 x 2048 step of the fp32 CPU model needs '
                                                         '>100 GB of attention matrices and ~1 minute; tokens/s is '
                                                         'per token and does not depend on the batch',
                         'dtype': 'f32 (the reference computes in fp32)', 'host_threads': threads},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'generate': {'metric': 'generated events/s (KV-cache decode, temperature 1.0)', 'value': gen_value,
                     'unit': 'events/s',
                     'cpu_baseline': {'value': gen_value, 'unit': 'events/s', 'cores': threads, 'kind': 'port',
                                      'sample': 'oracle cached decode (past=), fp32, 4 sequences x 256 events, '
                                                '%.1f s' % gen_seconds}},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------

def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from composer_b200 import _lib
    from composer_b200.models.transformer import Transformer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device; the product path has no CPU fallback '
                         '(use --impl reference for the CPU baseline).')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    if args.gpus != world and rank == 0:
        print('note: --gpus %d but WORLD_SIZE is %d; using WORLD_SIZE' % (args.gpus, world), file=sys.stderr)

    B, T, K, W = args.batch, SEQ_LEN, args.steps, max(args.warmup, 3)
    model = Transformer(VOCAB, MODEL['embedding_size'], MODEL['window_size'], MODEL['decoder_layers_count'],
                        MODEL['attention_head_count'], False, 0.0, 0.02, MODEL['attention_dropout_rate'],
                        MODEL['residual_dropout_rate'], 1e-5, True, True, device=device, seed=0)
    model.compile(1e-3)

    # synthetic batches: ids uniform on [0, vocab), labels = ids shifted by one (models/__init__.py:304)
    rng = np.random.default_rng(1234 + rank)
    n_batches = 4
    host_batches = []
    for _ in range(n_batches):
        draw = torch.from_numpy(rng.integers(0, VOCAB, size=(B, T + 1)).astype(np.int32))
        host_batches.append((draw[:, :-1].contiguous().pin_memory(), draw[:, 1:].contiguous().pin_memory()))
    device_batches = [(x.to(device), y.to(device)) for x, y in host_batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for i in range(steps):
            step_fn(i)
        end.record()
        barrier()
        seconds = torch.tensor([start.elapsed_time(end) * 1e-3], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(seconds, op=dist.ReduceOp.MAX)
        return float(seconds)

    def resident_step(i):
        x, y = device_batches[i % n_batches]
        model.train_step(x, y)

    losses = []

    def e2e_step(i):
        x, y = host_batches[i % n_batches]         # pinned host memory -> device inside the timed region
        loss_sum, _ = model.train_step(x, y)
        losses.append(float(loss_sum))              # device -> host read of the step's loss

    for i in range(W):
        resident_step(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches_before = _lib.call('cb200_launch_count')
    seconds = timed(resident_step, K)
    launches = _lib.call('cb200_launch_count') - launches_before
    clocks = sampler.stop() if rank == 0 else None
    for i in range(2):
        e2e_step(i)
    losses.clear()
    e2e_seconds = timed(e2e_step, K)

    tokens_per_step = world * B * T
    value = tokens_per_step * K / seconds
    e2e_value = tokens_per_step * K / e2e_seconds
    flops = step_flops_per_token(MODEL['decoder_layers_count'], MODEL['embedding_size'], VOCAB, T)

    # ---- roofline of the dominant kernel (attention backward), timed alone on this stream ----
    peaks, peak_kind = load_peaks()
    roofline, kernel_breakdown = None, None
    if rank == 0:
        roofline, kernel_breakdown = dominant_kernel_roofline(model, B, T, peaks, peak_kind, seconds / K, clocks)

    # ---- generation (configs[2]): reported beside the headline, same run ----
    generate = None
    if args.generate:
        generate = generation_benchmark(model, world, rank, device, dist if world > 1 else None,
                                        cpu_baseline=not args.no_cpu_baseline)

    # ---- the other single-GPU configurations of BASELINE.json, short runs (N = 1 only) ----
    other = None
    if world == 1 and args.other_configs:
        del model
        torch.cuda.empty_cache()
        other = other_configs(peaks)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_steps = 6
        cpu_value, threads, cpu_seconds = cpu_reference_step(1, T, cpu_steps, 1)
        cpu = {'value': cpu_value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': 'oracle training step (fp32 torch-CPU restatement of the reference, dropout 0.1; TensorFlow '
                         'not installable), 1 of the %d sequences x T=%d per step, 1 warm-up + %d timed steps of %.1f s'
                         % (B, T, cpu_steps, cpu_seconds)}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': seconds / K * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16', 'data': 'synthetic',
            'config': gpu_config(world, B),
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': 2 * B * T * 4,
                    'd2h_bytes_per_step': 4, 'ms_per_step': e2e_seconds / K * 1e3},
            'gpu_launches': int(launches),
            'step_tflops_per_gpu': value / world * flops / 1e12,
            'step_frac_of_sustained_bf16_peak': value / world * flops / 1e12 / peaks['bf16_tflops_sustained'],
            'roofline': roofline,
            'kernels': kernel_breakdown,
            'cpu_baseline': cpu,
            'generate': generate,
            'other_configs': other,
            'last_loss': losses[-1] / (B * T) if losses else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def dominant_kernel_roofline(model, B, T, peaks, peak_kind, step_seconds, clocks=None):
    '''
    Times the kernels of one decoder block alone (CUDA events on the launch
    stream) and returns the roofline entry of the one that dominates the step,
    plus the per-kernel table.  Algorithmic FLOPs of causal attention backward:
    2.5 x forward = 2.5 x 4 d_h T(T+1)/2 per (sequence, head).
    '''

    import ctypes
    import math

    import torch

    from composer_b200 import _lib

    E, H, L = model.embedding_size, model.attention_head_count, model.decoder_layers_count
    D = E // H
    dev = model.device

    def ptr(t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else None

    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def timeit(fn, iters=5):
        fn()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(iters):
            fn()
        end.record()
        torch.cuda.synchronize()
        return start.elapsed_time(end) * 1e-3 / iters

    scale = 1.0 / math.sqrt(D)
    qkv = torch.randn(B, T, 3 * E, device=dev).to(torch.bfloat16)
    out = torch.empty(B, T, E, device=dev, dtype=torch.bfloat16)
    dout = torch.randn(B, T, E, device=dev).to(torch.bfloat16)
    lse = torch.empty(B, H, T, device=dev)
    delta = torch.empty(B, H, T, device=dev)
    dq_acc = torch.zeros(B, T, E, device=dev)
    dqkv = torch.empty(B, T, 3 * E, device=dev, dtype=torch.bfloat16)
    rate = model.attention_dropout_rate
    t_fwd = timeit(lambda: _lib.call('cb200_attention_fwd', ptr(qkv), ptr(out), ptr(lse), B, T, H, D, scale, rate, 1,
                                     1, 1, stream))
    t_bwd = timeit(lambda: _lib.call('cb200_attention_bwd', ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(delta),
                                     ptr(dq_acc), ptr(dqkv), B, T, H, D, scale, rate, 1, 1, 1, stream))
    pairs = B * H * T * (T + 1) / 2.0
    fwd_flops = 4.0 * D * pairs
    # backward = 2 x forward FLOPs, the recomputation of S = Q K^T inside the kernel is not
    # counted (`achieved`); `achieved_with_recompute` counts the 5 MMAs the kernel really issues (2.5 x forward).
    bwd_flops = 2.0 * fwd_flops
    traffic, traffic_note = None, 'no ncu capture of this source tree under profiles/'
    summary_path = os.path.join(ROOT, 'profiles', 'ncu_summary.json')
    if os.path.exists(summary_path):
        from composer_b200 import build as native_build
        with open(summary_path) as handle:
            summary = json.load(handle)
        for name, entry in summary.items():
            if name.startswith('attn_bwd_tc_kernel') and entry.get('source_hash') == native_build.source_hash():
                traffic = entry.get('dram_bytes_per_launch')
                traffic_note = 'ncu --set full capture %s of this source tree (hash %s)' % (
                    entry.get('capture', '?'), entry['source_hash'][:12])
    achieved = bwd_flops / t_bwd / 1e12
    sm_mhz = (clocks or {}).get('sm_mhz') or 1965.0
    exp_peak = 16 * 148 * sm_mhz * 1e6          # MUFU.EX2: 16 per clock per SM
    roofline = {
        'kernel': 'attn_bwd_tc_kernel<%d,1> (tcgen05 / TMEM; + delta and dq-store helpers), one decoder block' % D,
        'bound': 'tensor', 'achieved': achieved, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
        'frac': achieved / peaks['bf16_tflops'], 'traffic': traffic, 'traffic_source': traffic_note,
        'peak_source': peak_kind + ' (burst)',
        'flop_convention': 'SURVEY 8(d): 2 x forward, recompute not counted',
        'achieved_with_recompute': 2.5 * fwd_flops / t_bwd / 1e12,
        'exp_per_s': pairs / t_bwd, 'exp_peak_per_s': exp_peak, 'exp_frac': pairs / t_bwd / exp_peak,
        'note': 'at d_h = 16 a score element costs one exponential and ~6 other instructions against 80 MMA FLOPs: '
                'the kernel is measured against the MUFU.EX2 rate (16 / clk / SM x 148 SMs x the SM clock sampled '
                'during the run) as well as the tensor peak',
        'share_of_step': L * t_bwd / step_seconds,
    }
    breakdown = {
        'attention_fwd_ms_per_block': t_fwd * 1e3, 'attention_bwd_ms_per_block': t_bwd * 1e3,
        'attention_share_of_step': L * (t_fwd + t_bwd) / step_seconds,
        'attention_fwd_tflops': fwd_flops / t_fwd / 1e12, 'attention_bwd_tflops': achieved,
        'attention_fwd_exp_frac': pairs / t_fwd / exp_peak,
    }
    return roofline, breakdown


def other_configs(peaks):
    '''
    configs[0], [3] and [4] of BASELINE.json on this GPU, short runs (3 warm-up + 3 timed steps each), so that the
    numbers DESIGN.md quotes for them are driver-visible.  configs[1] is the headline above, configs[2] is `generate`.
    '''

    import numpy as np
    import torch

    from composer_b200.models.transformer import Transformer

    def run(name, layers, embedding, heads, T, B, train):
        model = Transformer(VOCAB, embedding, T, layers, heads, False, 0.0, 0.02, 0.1, 0.1, 1e-5, True, True, seed=0)
        model.compile(1e-3)
        rng = np.random.default_rng(0)
        draw = torch.from_numpy(rng.integers(0, VOCAB, size=(B, T + 1)).astype(np.int32)).cuda()
        x, y = draw[:, :-1].contiguous(), draw[:, 1:].contiguous()
        fn = (lambda: model.train_step(x, y)) if train else (lambda: model.forward_loss(x, y, training=False))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(3):
            fn()
        end.record()
        torch.cuda.synchronize()
        seconds = start.elapsed_time(end) * 1e-3 / 3
        flops = step_flops_per_token(layers, embedding, VOCAB, T) / (1 if train else 3)
        tokens_per_s = B * T / seconds
        del model
        torch.cuda.empty_cache()
        return {'config': name, 'layers': layers, 'embedding_size': embedding, 'heads': heads, 'seq_len': T,
                'per_gpu_batch': B, 'phase': 'train step' if train else 'forward + loss', 'dropout': 0.1 if train else 0.0,
                'tokens_per_s': tokens_per_s, 'ms_per_step': seconds * 1e3,
                'algorithmic_tflops': tokens_per_s * flops / 1e12,
                'frac_of_sustained_bf16_peak': tokens_per_s * flops / 1e12 / peaks['bf16_tflops_sustained']}

    return [run('configs[0] forward + loss, default hyperparameters, B 4 x T 1024', 8, 256, 16, 1024, 4, False),
            run('configs[3] long-context training step, T 4096, B 16', 8, 256, 16, 4096, 16, True),
            run('configs[4] scaled Transformer (12 layers, d_model 1024) training step, T 1024, B 16', 12, 1024, 16,
                1024, 16, True)]


def decode_implementation(model, batch):
    '''Which decode path cb200_generate takes for this batch (mirrors decode_mega's choice of cluster size).'''

    from composer_b200 import _lib

    cap8 = _lib.call('cb200_decode_cluster_capacity', model._engine, 8)
    cap4 = _lib.call('cb200_decode_cluster_capacity', model._engine, 4)
    if cap8 <= 0 and cap4 <= 0:
        return {'kernel': 'per-step kernels replayed as a CUDA graph'}
    size = 8 if cap8 > 0 and (batch <= cap8 * 8 or cap4 <= 0) else 4
    cap = cap8 if size == 8 else cap4
    clusters = min(batch, cap) if min(batch, cap) * 8 >= batch else -(-batch // 8)
    share = -(-batch // clusters)          # the fewest clusters with the same largest share of sequences
    clusters = -(-batch // share)
    return {'kernel': 'decode_mega_kernel (one persistent launch per generation)', 'cluster_size': size,
            'clusters': clusters, 'co_resident_clusters': cap, 'launches_per_generation': 3}


def generation_benchmark(model, world, rank, device, dist, cpu_baseline=True):
    '''configs[2]: 256 sequences sharded over the GPUs, prompt 1, 1024 events, temperature 1.0, KV cache.'''

    import numpy as np
    import torch

    from composer_b200.models.transformer import Transformer

    total, length = 256, 1024
    per_rank = total // world
    gen_model = Transformer(VOCAB, MODEL['embedding_size'], 1024, MODEL['decoder_layers_count'],
                            MODEL['attention_head_count'], False, 0.0, 0.02, 0.1, 0.1, 1e-5, True, True,
                            device=device, seed=0)
    rng = np.random.default_rng(99)
    prompt = rng.integers(0, VOCAB, size=(total, 1))[rank * per_rank:(rank + 1) * per_rank]
    # warm-up at the timed shape: the KV cache and workspace are allocated here, not inside the timed region
    gen_model.generate(prompt, length, temperature=1.0, seed=7, sequence_index_base=rank * per_rank)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    # two timed generations (each: barrier, wall clock around the call, max over ranks); the faster one is reported
    # and both are listed: a whole generation is one kernel launch, a single timing has nothing to average over
    runs = []
    for _ in range(2):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        start = time.perf_counter()
        out = gen_model.generate(prompt, length, temperature=1.0, seed=7, sequence_index_base=rank * per_rank)
        host_ids = out.cpu()                        # the ids leave the device inside the timed region
        elapsed = torch.tensor([time.perf_counter() - start], device=device, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        runs.append(float(elapsed))
    seconds = min(runs)
    E, L = MODEL['embedding_size'], MODEL['decoder_layers_count']
    weights = 2 * (gen_model.count_params() - 1024 * E + E)
    kv_read = sum(per_rank * 2 * L * E * 2 * t for t in range(length))
    kv_write = per_rank * 2 * L * E * 2 * length
    bytes_per_gpu = weights * length + kv_read + kv_write
    peaks, kind = load_peaks()
    # DRAM bytes of one generation at this shape (one launch of the persistent kernel) from the committed ncu capture
    traffic = None
    summary_path = os.path.join(ROOT, 'profiles', 'ncu_summary.json')
    if world == 1 and os.path.exists(summary_path):
        from composer_b200 import build as native_build
        with open(summary_path) as handle:
            summary = json.load(handle)
        for name, entry in summary.items():      # decode_mega_kernel<d_h, cluster size, profiled>
            if name.startswith('decode_mega_kernel<16, 4') and 'prompt 1, 1,024 events' in entry.get('capture', '') \
                    and entry.get('source_hash') == native_build.source_hash():
                traffic = entry.get('dram_bytes_per_launch')
    cpu = None
    if rank == 0 and world == 1 and cpu_baseline:
        cpu_value, threads, cpu_seconds = cpu_reference_generate(4, 256)
        cpu = {'value': cpu_value, 'unit': 'events/s', 'cores': threads, 'kind': 'port',
               'sample': 'oracle cached decode (past=), fp32, 4 sequences x 256 events (%.1f s)' % cpu_seconds}
    return {'metric': 'generated events/s (256 sequences x 1024 events, temperature 1.0, KV-cache decode)',
            'value': total * length / seconds, 'unit': 'events/s', 'seconds': seconds, 'seconds_of_each_run': runs,
            'd2h_bytes_per_generation': int(host_ids.numel() * host_ids.element_size()), 'cpu_baseline': cpu,
            'sequences_per_gpu': per_rank,
            'us_per_step': seconds / length * 1e6,
            'roofline': {'bound': 'hbm', 'achieved': bytes_per_gpu / seconds / 1e9, 'peak': peaks['hbm_gbs'],
                         'unit': 'GB/s', 'frac': bytes_per_gpu / seconds / 1e9 / peaks['hbm_gbs'], 'traffic': traffic,
                         'algorithmic_bytes': bytes_per_gpu,
                         'peak_source': kind},
            'sample_ids': host_ids[0, :8].tolist(),
            'implementation': decode_implementation(gen_model, per_rank)}


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=10)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    parser.add_argument('--batch', type=int, default=BATCH_PER_GPU, help='sequences per GPU')
    parser.add_argument('--no-generate', dest='generate', action='store_false')
    parser.add_argument('--no-cpu-baseline', action='store_true')
    parser.add_argument('--no-other-configs', dest='other_configs', action='store_false')
    args = parser.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
