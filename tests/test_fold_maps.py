"""CPU checks of the index maps behind the TMEM-operand kernels (volpick_b200/csrc/fused_dec2.cu, fused_enc.cu).

The kernels hold, in TMEM lane r, everything that descends from row R0 + r of the 375-sample level; a conv tap is a COLUMN
offset ("slot") of the lane's operand.  These tests restate the host-side weight builders (decb2_build, enca_build,
tc_build_layer's polyphase form) in NumPy and compare the resulting banded / block GEMMs with the plain SeisBench ops
(nn.Upsample(x2) + conv1d 'same' + ReLU; conv1d 'same' + ReLU + MaxPool1d(2)) on random data -- no GPU, no library call.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _fdiv2(a):
    return a // 2  # Python floors


def _polyphase(W):
    """tc_build_layer(TC_POLYPHASE): x2 nearest up-sampling folded into the weights.
    out[2 s + phi] = sum_j weff[phi][j] x[s + o_min + j]."""
    cout, cin, k = W.shape
    p = k // 2
    o_min, o_max = _fdiv2(0 - p), _fdiv2(1 + (k - 1) - p)
    ntaps = o_max - o_min + 1
    weff = np.zeros((2, ntaps, cout, cin))
    for phi in range(2):
        for kk in range(k):
            weff[phi, _fdiv2(phi + kk - p) - o_min] += W[:, :, kk]
    return weff, o_min, ntaps


def _ref_up_conv(x, W):
    """nn.Upsample(scale_factor=2, mode='nearest') + conv1d(padding=k // 2): x (cin, T) -> (cout, 2 T)."""
    xu = torch.from_numpy(x).repeat_interleave(2, dim=1)[None]
    return F.conv1d(xu, torch.from_numpy(W), padding=W.shape[2] // 2)[0].numpy()


def test_polyphase_blocks_of_the_tmem_layers():
    """decoder.convs.5 / .6 in fused_dec2.cu: block b of lane R = input sample f R + b -> outputs 2 (f R + b) + phi, tap j reads
    slot b + j where slot i holds input sample f R - halo + i (f = 4, halo 2 for convs.5; f = 8, halo 3 for convs.6)."""
    rng = np.random.default_rng(0)
    for (cin, cout, k, f, halo, taps) in [(16, 16, 9, 4, 2, 5), (16, 8, 11, 8, 3, 7)]:
        W = rng.standard_normal((cout, cin, k))
        weff, o_min, ntaps = _polyphase(W)
        assert (o_min, ntaps) == (-halo, taps)
        rows = 12
        x = rng.standard_normal((cin, f * rows))
        ref = _ref_up_conv(x, W)
        xp = np.pad(x, ((0, 0), (halo, halo)))  # rows outside the sequence are zeros at every level
        for R in range(rows):
            slots = xp[:, f * R: f * R + f + 2 * halo]  # slot i = sample f R - halo + i
            for b in range(f):
                for phi in range(2):
                    acc = sum(weff[phi, j] @ slots[:, b + j] for j in range(ntaps))
                    np.testing.assert_allclose(acc, ref[:, 2 * (f * R + b) + phi], atol=1e-10)


def test_decoder_convs4_folded_over_two_samples():
    """decb2_build: decoder.convs.4 (16, 32, 7) after x2 up-sampling, on rows that hold the two 750-level samples of a 375-level
    row: three row taps (R - 1, R, R + 1), K index = (sample in row) * 32 + ci, column n = s' * 16 + co (s' = 0..3)."""
    rng = np.random.default_rng(1)
    W = rng.standard_normal((16, 32, 7))
    wf = np.zeros((3, 64, 64))
    for sp in range(4):
        for kk in range(7):
            u = sp + kk - 3          # up-sampled position relative to 4 R
            d = _fdiv2(u)            # 750-level position relative to 2 R
            jr, ph = _fdiv2(d) + 1, d - 2 * _fdiv2(d)
            assert 0 <= jr < 3
            for co in range(16):
                wf[jr, ph * 32: ph * 32 + 32, sp * 16 + co] += W[co, :, kk]
    rows = 10
    x = rng.standard_normal((32, 2 * rows))  # 750 level
    ref = _ref_up_conv(x, W)                 # (16, 4 rows)
    xr = np.pad(x, ((0, 0), (2, 2))).T.reshape(rows + 2, 64)  # row i = 375-level row i - 1: [sample 0 | sample 1] x 32 channels
    for R in range(rows):
        out = sum(xr[R + jr] @ wf[jr] for jr in range(3))
        np.testing.assert_allclose(out.reshape(4, 16).T, ref[:, 4 * R: 4 * R + 4], atol=1e-10)


def test_encoder_convs1_folded_over_eight_samples_and_the_lane_local_pool():
    """enca_build: encoder.convs.1 (16, 8, 9) 'same' on rows of eight 3000-level samples: K steps = (row tap jr, plane pair p0),
    input sample d = (jr - 1) * 8 + p0 + k-half relative to 8 R, weight tap k = d - t + 4; then Conv + ReLU + MaxPool1d(2) =
    relu(max(column t, column t + 1) + bias) inside the lane."""
    rng = np.random.default_rng(2)
    W, bias = rng.standard_normal((16, 8, 9)), rng.standard_normal(16)
    JR, P0 = [0, 0, 1, 1, 1, 1, 2, 2], [4, 6, 0, 2, 4, 6, 0, 2]
    rows = 9
    x = rng.standard_normal((8, 8 * rows))
    conv = F.conv1d(torch.from_numpy(x)[None], torch.from_numpy(W), torch.from_numpy(bias), padding=4)
    ref = F.max_pool1d(F.relu(conv), 2)[0].numpy()  # (16, 4 rows)
    xr = np.pad(x, ((0, 0), (8, 8))).T.reshape(rows + 2, 8, 8)  # [row + 1][sample in row][channel]
    for R in range(rows):
        acc = np.zeros((8, 16))  # [conv sample t][co]
        for ks in range(8):
            for kh in range(2):
                d = (JR[ks] - 1) * 8 + P0[ks] + kh
                xin = xr[R + JR[ks], P0[ks] + kh]  # 8 channels of sample 8 R + d
                for t in range(8):
                    k = d - t + 4
                    if 0 <= k <= 8:
                        acc[t] += W[:, :, k] @ xin
        pooled = np.maximum(np.maximum(acc[0::2], acc[1::2]) + bias, 0.0)  # (4, 16)
        np.testing.assert_allclose(pooled.T, ref[:, 4 * R: 4 * R + 4], atol=1e-10)


def test_encoder_tmem_slots_and_two_lane_halo():
    """fused_enc.cu: convs.2 (k = 7) on lanes of four 1500-level samples (10 slots = samples 4R-3 .. 4R+6, conv sample t reads
    slots t .. t + 6) and convs.3 on lanes of two 750-level samples (8 slots = samples 2R-3 .. 2R+4: the halo comes from lanes
    R +- 1 AND R +- 2)."""
    rng = np.random.default_rng(3)
    for f, halo_lanes in [(4, 1), (2, 2)]:
        W = rng.standard_normal((16, 16, 7))
        rows = 9
        x = rng.standard_normal((16, f * rows))
        ref = F.conv1d(torch.from_numpy(x)[None], torch.from_numpy(W), padding=3)[0].numpy()
        lanes = np.pad(x, ((0, 0), (f * halo_lanes, f * halo_lanes))).T.reshape(rows + 2 * halo_lanes, f, 16)  # [lane][sample][ch]
        for R in range(rows):
            # the slots of lane R assembled from its own samples and its neighbours' (what the epilogue's shuffles copy)
            slot = []
            for i in range(f + 6):
                s = f * R - 3 + i
                lane, off = divmod(s, f)
                assert abs(lane - R) <= halo_lanes
                slot.append(lanes[lane + halo_lanes, off])
            for t in range(f):
                acc = sum(W[:, :, j] @ slot[t + j] for j in range(7))
                np.testing.assert_allclose(acc, ref[:, f * R + t], atol=1e-10)


def test_item_tiling_covers_the_kept_rows():
    """decb2_launch / enca_launch: items of 112 (94) used rows cover the kept rows of a window; lanes outside are never stored."""
    for keep_lo, keep_hi in [(0, 6000), (500, 5500), (1000, 5000), (0, 16), (5984, 6000)]:
        row_off0, row_hi = keep_lo // 16, (keep_hi + 15) // 16
        tiles = (row_hi - row_off0 + 111) // 112
        covered = set()
        for j in range(tiles):
            R0 = row_off0 + 112 * j - 4
            covered |= {R0 + r for r in range(4, 116) if 0 <= R0 + r < 375 and R0 + r < row_hi}
        assert covered == set(range(row_off0, min(row_hi, 375)))
        if (keep_lo, keep_hi) == (500, 5500):
            assert tiles == 3  # a blinded window: three items, as with the 120 valid lanes
    tiles = (375 + 93) // 94
    assert {94 * j - 3 + r for j in range(tiles) for r in range(3, 97) if 94 * j - 3 + r < 375} == set(range(375))
