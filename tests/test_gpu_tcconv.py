"""GPU parity of the tcgen05 implicit-GEMM Conv1d layer kernel against torch.nn.functional on the CPU."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from volpick_b200 import _lib

pytestmark = pytest.mark.gpu

CASES = [
    # name, NS, CIN, T, COUT, K, mode, crop, act, pool
    ("enc2_direct_pool", 5, 16, 1500, 16, 7, 0, 0, 1, 2),
    ("enc4_odd_pool", 7, 32, 375, 32, 5, 0, 0, 1, 2),
    ("enc6_direct_pool", 9, 64, 94, 64, 3, 0, 0, 1, 2),
    ("enc1_cin8_tapsk", 3, 8, 3000, 16, 9, 0, 0, 1, 2),
    ("head_cin8_sigmoid", 2, 8, 6000, 1, 11, 0, 0, 2, 1),
    ("res_k3_t47", 11, 64, 47, 64, 3, 0, 0, 0, 1),
    ("dec0_poly_k3", 13, 16, 47, 64, 3, 1, 0, 1, 1),
    ("dec1_poly_k5", 6, 64, 94, 64, 5, 1, 0, 1, 1),
    ("dec2_ups_crop", 6, 64, 188, 32, 5, 2, 1, 1, 1),
    ("dec3_poly_k7", 4, 32, 375, 32, 7, 1, 0, 1, 1),
    ("dec4_poly_k7", 3, 32, 750, 16, 7, 1, 0, 1, 1),
    ("dec5_poly_k9", 2, 16, 1500, 16, 9, 1, 0, 1, 1),
    ("dec6_poly_k11", 2, 16, 3000, 8, 11, 1, 0, 1, 1),
    ("cin3_padded", 2, 3, 700, 8, 11, 0, 0, 1, 2),
    ("single_row_tile", 1, 16, 5, 16, 7, 0, 0, 0, 1),
    # mode 3: the same conv on the [T / 4][4 C] view, ReLU + MaxPool1d(2) inside the accumulator row (encoder.convs.1 / .2)
    ("enc1_fold4_pool", 3, 8, 3000, 16, 9, 3, 0, 1, 2),
    ("enc2_fold4_pool", 5, 16, 1500, 16, 7, 3, 0, 1, 2),
    ("fold4_pool_short", 2, 16, 8, 16, 7, 3, 0, 1, 2),
]


def _reference(x, w, b, mode, crop, act, pool):
    xt = torch.from_numpy(x)
    if mode in (1, 2):
        xt = F.interpolate(xt, scale_factor=2, mode="nearest")
        if crop:
            xt = xt[:, :, :-crop]
    y = F.conv1d(xt, torch.from_numpy(w), torch.from_numpy(b), padding=w.shape[2] // 2)
    if act == 1:
        y = torch.relu(y)
    if act == 2:
        y = torch.sigmoid(y)
    if pool == 2:
        if y.shape[2] % 2:
            y = F.pad(y, (0, 1), "constant", -1e10)
        y = F.max_pool1d(y, 2)
    return y.numpy()


@pytest.mark.parametrize("precision,tol", [("f16x3", 2e-5), ("bf16", 6e-2)])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tcconv_layer(case, precision, tol):
    name, NS, CIN, T, COUT, K, mode, crop, act, pool = case
    lib = _lib.load()
    rng = np.random.default_rng(abs(hash(name)) % 2**31)
    x = (rng.standard_normal((NS, CIN, T)) * 2).astype(np.float32)
    if act != 2:
        x = np.abs(x)  # post-ReLU-like activations, as inside the network
    w = (rng.standard_normal((COUT, CIN, K)) / np.sqrt(CIN * K)).astype(np.float32)
    b = rng.standard_normal(COUT).astype(np.float32) * 0.1
    ref = _reference(x, w, b, mode, crop, act, pool)
    d_x = torch.from_numpy(x).cuda()
    d_y = torch.full(ref.shape, float("nan"), dtype=torch.float32, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.vp_tcconv_debug(d_x.data_ptr(), NS, CIN, T, w.ctypes.data, b.ctypes.data, COUT, K, mode, crop, act, pool,
                                   _lib.PRECISION[precision], d_y.data_ptr(), stream))
    got = d_y.cpu().numpy()
    assert not np.isnan(got).any(), f"{name}: {np.isnan(got).sum()} outputs never written"
    err = float(np.abs(got - ref).max())
    scale = float(np.abs(ref).max())
    print(f"{name} {precision}: max|diff| = {err:.3e} (max|ref| = {scale:.3e})")
    assert err <= tol * max(1.0, scale), (name, precision, err, scale)
