"""The C-ABI library builds, loads and exports every symbol include/smplify_b200.h declares.
No compute calls (CPU-only box)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from inbed_pose_estimation_b200 import _native


def _declared():
    text = open(os.path.join(ROOT, 'include', 'smplify_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(smplb200_[a-z_0-9]+)\s*\(', text)))


def test_library_builds_and_exports_header_symbols():
    path = _native.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), 'missing export ' + n
    assert sorted(_native.EXPORTED_SYMBOLS) == names
    assert _native.lib().smplb200_version() == 200


def test_model_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from inbed_pose_estimation_b200 import synthetic
    arrays = synthetic.model_arrays(0)
    arrays['J_regressor_extra'] = synthetic.make_extra_regressor(1)
    with pytest.raises(RuntimeError, match='no CUDA device|CPU fallback'):
        _native.NativeModel(arrays, None, 0)


def test_workspace_queries():
    lib = _native.lib()
    assert lib.smplb200_fit_workspace_bytes(0) >= 0
    assert lib.smplb200_fit_workspace_bytes(4096) >= 4096 * 512 * 4
    assert lib.smplb200_smpl_workspace_bytes(32) >= 32 * 512 * 4 * 9


def test_host_side_validation():
    """Python wrappers reject CPU tensors and bad shapes before reaching the device."""
    import torch
    from inbed_pose_estimation_b200 import geometry
    with pytest.raises(RuntimeError, match='CUDA'):
        geometry.batch_rodrigues(torch.zeros(3, 3))
    with pytest.raises(RuntimeError, match='CUDA'):
        geometry.perspective_projection(torch.zeros(1, 2, 3), torch.eye(3)[None], torch.zeros(1, 3), 5000., torch.zeros(1, 2))


def test_fit_tile_plan_covers_every_batch_in_the_fewest_waves():
    """smplb200_fit_tile_plan: host arithmetic behind the fit launch (no GPU).  Every sample is covered, the 16-sample tiles
    form whole waves, the remainder is one wave of the smallest tile size that holds it."""
    sms = 148
    for batch in list(range(1, 70)) + [255, 256, 592, 593, 1184, 1185, 1776, 1777, 2368, 2369, 4096, 4103, 4736, 65536, 100000]:
        n16, small, n_small = _native.fit_tile_plan(batch, sms)
        assert n16 * 16 + n_small * small >= batch
        assert n16 * 16 + max(n_small - 1, 0) * small < batch or n_small == 0 and (n16 - 1) * 16 < batch     # no empty tile
        assert small in (0, 4, 8, 12) and (n_small == 0) == (small == 0)
        assert 0 <= n_small <= sms
        if n_small:
            assert n16 % sms == 0                                            # whole waves of 16s before the small tiles
            rest = batch - n16 * 16
            assert all((rest + s - 1) // s > sms for s in (4, 8, 12) if s < small)      # no smaller size would fit one wave
    assert _native.fit_tile_plan(4096, sms) == (148, 12, 144)
    assert _native.fit_tile_plan(32, sms) == (0, 4, 8)
    assert _native.fit_tile_plan(4736, sms) == (296, 0, 0)
    assert _native.fit_tile_plan(0, sms) == (0, 0, 0)


def test_fit_split_plan_cluster_sizes():
    """smplb200_fit_split_plan: small batches run as clusters of 8 / 4 / 2 CTAs per 4-sample tile, the largest size whose
    clusters all fit the chip at once; beyond sms / 2 tiles the tile kernels take over (no GPU: host arithmetic only)."""
    sms = 148
    for batch in range(1, 700):
        c = _native.fit_split_plan(batch, sms)
        tiles = (batch + 3) // 4
        assert c in (0, 2, 4, 8)
        if c:
            assert tiles * c <= sms and (c == 8 or tiles * 2 * c > sms)
        else:
            assert tiles * 2 > sms
    assert [_native.fit_split_plan(b, sms) for b in (1, 32, 72, 73, 148, 149, 256, 296, 297, 4096)] == [8, 8, 8, 4, 4, 2, 2, 2, 0, 0]
    assert _native.fit_split_plan(0, sms) == 0 and _native.fit_split_plan(32, 1) == 0


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: no file of the package (Python or CUDA/C++) may import, include or open it."""
    pkg = os.path.join(ROOT, 'inbed_pose_estimation_b200')
    for folder, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                continue
            text = open(os.path.join(folder, f), errors='replace').read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
            assert not re.search(r'#\s*include\s*[<"][^>"]*oracle', text), f
            for literal in re.findall(r'"[^"\n]*"|\'[^\'\n]*\'', text):           # paths only appear in comments
                assert 'oracle' not in literal and '/root/reference' not in literal, (f, literal)
