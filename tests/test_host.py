'''
Host-side checks that need no GPU: the C ABI library loads and exports every
symbol include/composer_b200.h declares, the parameter arena layout matches
the oracle's Keras-ordered variables, the configuration / CLI helpers mirror
the reference, and the data-parallel helpers work over gloo with 2 processes.
'''

import ctypes
import os
import re

import numpy as np
import pytest

from composer_b200 import _lib, config as config_module, parallel
from oracle import transformer_oracle as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    with open(os.path.join(ROOT, 'include', 'composer_b200.h')) as handle:
        text = handle.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(cb200_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    library = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(library, name), 'libcomposer_b200.so does not export %s' % name
    # and the ctypes table covers the header
    assert set(declared) == set(_lib.exported_symbols())
    assert library.cb200_abi_version() == 1


def _default_config():
    return _lib.Config(390, 256, 1024, 8, 16, 0.1, 0.1, 1e-5, 1, 1)


def test_parameter_layout_matches_keras_order():
    cfg = _default_config()
    count = _lib.call('cb200_param_tensor_count', ctypes.byref(cfg))
    shapes = oracle.parameter_shapes(oracle.OracleConfig())
    assert count == len(shapes)
    name = ctypes.create_string_buffer(128)
    offset, rows, cols = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32()
    cursor = 0
    for index, (expected_name, shape) in enumerate(shapes.items()):
        _lib.call('cb200_param_tensor_info', ctypes.byref(cfg), index, name, 128, ctypes.byref(offset),
                  ctypes.byref(rows), ctypes.byref(cols))
        assert name.value.decode() == expected_name
        assert rows.value * cols.value == int(np.prod(shape))
        assert offset.value == cursor and offset.value % 4 == 0
        cursor += rows.value * cols.value
    assert _lib.call('cb200_param_elems', ctypes.byref(cfg)) == 6680576


def test_unsupported_configurations_are_rejected_with_a_message():
    bad = _lib.Config(390, 250, 1024, 8, 16, 0.1, 0.1, 1e-5, 1, 1)     # 250 not divisible by 16
    with pytest.raises(_lib.NativeError) as error:
        _lib.call('cb200_param_tensor_info', ctypes.byref(bad), 0, None, 0, None, None, None)
    assert 'divisible' in str(error.value)
    huge_vocab = _lib.Config(5000, 256, 1024, 8, 16, 0.1, 0.1, 1e-5, 1, 1)
    assert _lib.load().cb200_param_elems(ctypes.byref(huge_vocab)) == -1
    assert b'512' in _lib.load().cb200_last_error()


def test_default_config_file_matches_reference_values():
    cfg = config_module.get(os.path.join(ROOT, 'composer_b200', 'default_config.yml'))
    model = cfg.transformer.model
    assert (model.window_size, model.embedding_size, model.decoder_layers_count, model.attention_head_count) == \
        (1024, 256, 8, 16)
    assert (model.attention_dropout_rate, model.residual_dropout_rate) == (0.1, 0.1)
    assert model.layer_normalization_epsilon == 1e-5 and model.scale_attention and model.use_layer_normalization
    assert (cfg.transformer.train.batch_size, cfg.transformer.train.learning_rate) == (1, 0.001)
    assert (cfg.dataset.time_step_increment, cfg.dataset.max_time_steps, cfg.dataset.velocity_bins) == (10, 100, 32)
    reference = '/root/reference/composer/default_config.yml'
    if os.path.exists(reference):
        theirs = config_module.get(reference)
        assert dict(theirs.transformer.model) == dict(model)
        assert dict(theirs.transformer.train) == dict(cfg.transformer.train)
        assert {k: v for k, v in theirs.dataset.items()} == {k: v for k, v in cfg.dataset.items()}


def test_gradient_buckets_partition_the_arena():
    cfg = _default_config()
    layout = {}
    name = ctypes.create_string_buffer(128)
    offset, rows, cols = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32()
    for index in range(_lib.call('cb200_param_tensor_count', ctypes.byref(cfg))):
        _lib.call('cb200_param_tensor_info', ctypes.byref(cfg), index, name, 128, ctypes.byref(offset),
                  ctypes.byref(rows), ctypes.byref(cols))
        layout[name.value.decode()] = (offset.value, rows.value, cols.value)
    buckets = parallel.gradient_buckets(layout, 8)
    assert len(buckets) == 10
    covered = sorted(buckets)
    assert covered[0][0] == 0 and covered[-1][1] == 6680576
    for (a0, a1), (b0, b1) in zip(covered, covered[1:]):
        assert a1 == b0
    # backward order: ln_f first, block 8 next, embeddings last
    assert buckets[0][0] == layout['ln_f/gamma'][0]
    assert buckets[1][0] == layout['h_8/ln_1/gamma'][0]
    assert buckets[-1][0] == 0


def test_shard_range_covers_everything_once():
    for total, world in ((256, 8), (10, 4), (3, 8), (0, 2)):
        seen = []
        for rank in range(world):
            start, end = parallel.shard_range(total, world, rank)
            seen.extend(range(start, end))
        assert seen == list(range(total))


def _gloo_worker(rank, world, port, results):
    import torch
    import torch.distributed as dist

    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    layout = {'wte/weight': (0, 6, 4), 'wpe/embeddings': (24, 5, 4),
              'h_1/ln_1/gamma': (44, 1, 4), 'h_1/mlp/c_proj/bias': (48, 1, 4),
              'h_2/ln_1/gamma': (52, 1, 4), 'h_2/mlp/c_proj/bias': (56, 1, 4),
              'ln_f/gamma': (60, 1, 4), 'ln_f/beta': (64, 1, 4)}
    buckets = parallel.gradient_buckets(layout, 2)
    flat = torch.arange(68, dtype=torch.float32) * (rank + 1)
    order = []
    parallel.allreduce_buckets(flat, buckets, after_bucket=order.append)
    expected = torch.arange(68, dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok = bool(torch.equal(flat, expected)) and order == [0, 1, 2, 3]
    # the mean of per-rank means over equal shards equals the global mean (what folding 1/world into Adam relies on)
    shard = torch.full((4,), float(rank))
    total = shard.mean().clone()
    dist.all_reduce(total)
    ok = ok and abs(float(total) / world - (world - 1) / 2) < 1e-6
    # without --seed every process draws its own seed from the clock: all ranks must adopt rank 0's
    ok = ok and parallel.agree_on_seed(1000 + 17 * rank) == 1000
    # ... while the dropout key differs between ranks (and is the plain seed on rank 0)
    keys = {parallel.rank_dropout_seed(1000, r) for r in range(world)}
    ok = ok and len(keys) == world and parallel.rank_dropout_seed(1000, 0) == 1000
    results[rank] = ok
    dist.destroy_process_group()


def test_bucketed_allreduce_over_gloo_world_size_2():
    import torch.multiprocessing as mp

    manager = mp.Manager()
    results = manager.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}


def test_tools_and_bench_compile():
    '''The measurement scripts are not imported by any other test (most need a GPU, one needs TensorFlow): at
    least their syntax is checked here, so that an edit cannot leave the driver's bench or a profiling recipe
    unparsable.'''
    import glob
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    scripts = sorted(glob.glob(os.path.join(root, 'tools', '*.py'))) + [os.path.join(root, 'bench.py'),
                                                                      os.path.join(root, '__graft_entry__.py')]
    assert len(scripts) >= 10
    for path in scripts:
        py_compile.compile(path, doraise=True)


def test_agree_on_seed_is_identity_without_a_process_group():
    assert parallel.agree_on_seed(12345) == 12345
    assert 0 <= parallel.rank_dropout_seed(-5, 3) < 2 ** 64


def test_prefetcher_propagates_errors_and_stops_its_thread():
    import threading
    import time

    import numpy as np

    from composer_b200 import data

    class Broken:
        def __len__(self):
            return 3

        def __iter__(self):
            yield np.zeros((2, 4), dtype=np.int64), np.zeros((2, 4), dtype=np.int64)
            raise OSError('bad file')

    iterator = iter(data.prefetch_pinned(Broken()))
    next(iterator)
    try:
        next(iterator)
        raised = False
    except OSError:
        raised = True
    assert raised, 'an exception in the loader thread must not look like the end of the epoch'

    class Endless:
        def __len__(self):
            return 10 ** 9

        def __iter__(self):
            while True:
                yield np.zeros((2, 4), dtype=np.int64), np.zeros((2, 4), dtype=np.int64)

    before = threading.active_count()
    iterator = iter(data.prefetch_pinned(Endless(), depth=2))
    next(iterator)
    iterator.close()                  # the consumer leaves early (max_steps): the producer must exit
    deadline = time.time() + 5
    while threading.active_count() > before and time.time() < deadline:
        time.sleep(0.05)
    assert threading.active_count() <= before
