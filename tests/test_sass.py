'''The Blackwell claims of DESIGN.md, checked in the SASS of the library that `build()` produced here (cuobjdump
needs no GPU): which kernels use tcgen05 / TMEM / TMA, that the decode kernel has a fence-free all-gather path, and
that the row-wise kernels compute on packed fp32 pairs.  `profiles/sass_opcodes_r2.txt` is the committed output of
the same tool.'''
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def table():
    if shutil.which('cuobjdump') is None or shutil.which('cu++filt') is None:
        pytest.skip('cuobjdump / cu++filt not on PATH')
    sys.path.insert(0, ROOT)
    from composer_b200 import build
    if not os.path.exists(build.LIBRARY) or not build.is_current():
        build.build(verbose=False)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'sass_opcodes.py')], capture_output=True, text=True,
                         check=True).stdout
    rows = {}
    for line in out.splitlines():
        if line.startswith('#') or not line.strip():
            continue
        name, rest = line[:88].rstrip(), line[88:].split()
        rows[name] = dict(item.split('=') for item in rest[1:])
    return rows


def _kernels(table, prefix):
    found = {name: ops for name, ops in table.items() if name.startswith(prefix)}
    assert found, 'no kernel named %s* in the library' % prefix
    return found


def test_tensor_core_kernels_are_tcgen05(table):
    # GEMMs: TMA loads and stores, tcgen05.mma, TMEM loads
    for name, ops in _kernels(table, 'gemm_sm100_kernel<').items():
        assert {'UTCHMMA', 'UTMALDG', 'LDTM'} <= set(ops), (name, ops)
    # training attention, forward and backward: tcgen05.mma on TMA-fed operands, scores read back from TMEM;
    # the default forward (bench configuration, dropout on and off) also writes the probabilities to TMEM for the
    # TS-form P V MMA (the A/B variants that pass P through shared memory do not)
    for name, ops in _kernels(table, 'attn_fwd_tc_kernel<').items():
        assert {'UTCHMMA', 'UTMALDG', 'LDTM'} <= set(ops), (name, ops)
    for drop in (0, 1):
        assert 'STTM' in table['attn_fwd_tc_kernel<16, %d, 0, 128, 2, 0, 1, 0>' % drop]
    for name, ops in _kernels(table, 'attn_bwd_tc_kernel<').items():
        assert {'UTCHMMA', 'UTMALDG', 'LDTM'} <= set(ops), (name, ops)


def test_decode_kernel_streams_by_tma_and_gathers_without_fences(table):
    kernels = _kernels(table, 'decode_mega_kernel<')
    for name, ops in kernels.items():
        assert 'UBLKCP' in ops and 'HMMA' in ops, (name, ops)           # TMA bulk copies, warp-level MMA
    # template arguments <d_h, cluster size, profile build, async gather>
    for name, ops in kernels.items():
        async_gather = name.rstrip('>').split(',')[-1].strip() == '1'
        barriers = int(ops.get('UCGABAR_ARV', 0))
        if async_gather:
            assert 'STAS' in ops and barriers == 3, (name, ops)          # start-up + the two per step
        else:
            assert 'STAS' not in ops and barriers == 4, (name, ops)      # + the shared one of the phase loop


def test_rowwise_kernels_use_packed_fp32(table):
    for prefix in ('layernorm_fwd_kernel<', 'layernorm_bwd_kernel<'):
        for name, ops in _kernels(table, prefix).items():
            assert 'FFMA2' in ops and 'FADD2' in ops, (name, ops)
