"""fp32 PyTorch-CPU restatement of the two network forwards (TEST INFRASTRUCTURE).

Follows SeisBench ``seisbench/models/eqtransformer.py`` (classes ``Encoder``,
``ResCNNBlock``, ``BiLSTMBlock``, ``Transformer``, ``SeqSelfAttention``,
``LayerNormalization``, ``FeedForward``, ``Decoder``, ``EQTransformer.forward``)
and ``seisbench/models/phasenet.py`` (``PhaseNet.__init__/forward/_merge_skip``)
as restated in SURVEY.md Appendix A / B.  The only reference-side pins are the
state-dict key names/shapes (/root/reference/Final_models/volpick/*/volpick.pt.v1),
the forward signature ``det, p, s = model(x)`` / ``(B,3,L)`` with labels "PSN"
(/root/reference/volpick/model/eval_taks0.py:68-72,85-89,131-134) and the
constructor calls (/root/reference/volpick/model/models.py:141,520).

Everything is driven directly from a ``state_dict`` with plain
``torch.nn.functional`` calls so that no hidden module default can leak in: every
constant that is not in the weight file is spelled out
below.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-3          # Appendix D #1 (Keras-compatible BatchNorm eps, both nets)
LN_EPS = 1e-14         # Appendix D #8
ATTN_EPS = 1e-5        # Appendix D #7 (original_compatible=False)
POOL_PAD_VALUE = -1e10  # Appendix D #4

SD = Dict[str, torch.Tensor]


def _bn(x: torch.Tensor, sd: SD, prefix: str) -> torch.Tensor:
    """BatchNorm1d in eval mode with running statistics, eps=1e-3."""
    return F.batch_norm(
        x,
        sd[prefix + ".running_mean"],
        sd[prefix + ".running_var"],
        sd[prefix + ".weight"],
        sd[prefix + ".bias"],
        training=False,
        eps=BN_EPS,
    )


def _lstm_direction(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh, reverse: bool) -> torch.Tensor:
    """One direction of a single-layer nn.LSTM; x is (T, B, C); gate order i,f,g,o; h0=c0=0."""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = x.new_zeros(T, B, H)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gates = x[t] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, g, o = gates.split(H, dim=1)
        i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
        c = f * c + i * g
        h = o * torch.tanh(c)
        out[t] = h
    return out


_LSTM_CACHE: Dict[Tuple[int, str], torch.nn.LSTM] = {}


def _nn_lstm(sd: SD, prefix: str, bidirectional: bool) -> torch.nn.LSTM:
    """The reference path uses a stock ``nn.LSTM(in, 16, bidirectional=...)``; build it from the state dict."""
    key = (id(sd), prefix)
    mod = _LSTM_CACHE.get(key)
    if mod is None:
        w_ih = sd[prefix + ".weight_ih_l0"]
        mod = torch.nn.LSTM(w_ih.shape[1], w_ih.shape[0] // 4, bidirectional=bidirectional)
        names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
        if bidirectional:
            names += [n + "_reverse" for n in names]
        mod.load_state_dict({n: sd[prefix + "." + n] for n in names})
        mod.eval()
        _LSTM_CACHE[key] = mod
    return mod


def lstm(x: torch.Tensor, sd: SD, prefix: str, bidirectional: bool, manual: bool = False) -> torch.Tensor:
    """nn.LSTM(in, 16, bidirectional=...) on (B, C, T) -> (B, 16*ndir, T).

    ``manual=True`` uses the explicit gate loop above instead of ``nn.LSTM`` (cross-check in tests).
    """
    xt = x.permute(2, 0, 1)  # (T, B, C)
    if not manual:
        with torch.no_grad():
            y = _nn_lstm(sd, prefix, bidirectional)(xt)[0]
        return y.permute(1, 2, 0).contiguous()
    outs = [
        _lstm_direction(
            xt,
            sd[prefix + ".weight_ih_l0"],
            sd[prefix + ".weight_hh_l0"],
            sd[prefix + ".bias_ih_l0"],
            sd[prefix + ".bias_hh_l0"],
            reverse=False,
        )
    ]
    if bidirectional:
        outs.append(
            _lstm_direction(
                xt,
                sd[prefix + ".weight_ih_l0_reverse"],
                sd[prefix + ".weight_hh_l0_reverse"],
                sd[prefix + ".bias_ih_l0_reverse"],
                sd[prefix + ".bias_hh_l0_reverse"],
                reverse=True,
            )
        )
    y = torch.cat(outs, dim=2)  # (T, B, H*ndir): forward units first, as nn.LSTM
    return y.permute(1, 2, 0).contiguous()


def seq_self_attention(x: torch.Tensor, sd: SD, prefix: str, width: Optional[int]) -> torch.Tensor:
    """Additive self attention (SeisBench ``SeqSelfAttention``), x (B, C, T) -> (B, C, T)."""
    xt = x.permute(0, 2, 1)  # (B, T, C)
    q = torch.unsqueeze(torch.matmul(xt, sd[prefix + ".Wt"]), 2)  # (B, T, 1, U)
    k = torch.unsqueeze(torch.matmul(xt, sd[prefix + ".Wx"]), 1)  # (B, 1, T, U)
    h = torch.tanh(q + k + sd[prefix + ".bh"])
    e = torch.squeeze(torch.matmul(h, sd[prefix + ".Wa"]) + sd[prefix + ".ba"], -1)  # (B, T, T)
    e = e - torch.max(e, dim=-1, keepdim=True).values
    e = torch.exp(e)
    if width is not None:
        T = e.shape[1]
        lower = torch.arange(0, T) - width // 2
        upper = lower + width
        indices = torch.unsqueeze(torch.arange(0, T), 1)
        mask = torch.logical_and(lower <= indices, indices < upper)
        e = torch.where(mask, e, torch.zeros_like(e))
    a = e / (torch.sum(e, dim=-1, keepdim=True) + ATTN_EPS)
    v = torch.matmul(a, xt)
    return v.permute(0, 2, 1).contiguous()


def layer_norm_channels(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """SeisBench ``LayerNormalization``: over the channel dim of (B, C, T), biased var, eps 1e-14."""
    mean = torch.mean(x, 1, keepdim=True)
    var = torch.mean((x - mean) ** 2, 1, keepdim=True) + LN_EPS
    return (x - mean) / torch.sqrt(var) * gamma + beta


def transformer(x: torch.Tensor, sd: SD, prefix: str) -> torch.Tensor:
    y = seq_self_attention(x, sd, prefix + ".attention", None)
    y = x + y
    y = layer_norm_channels(y, sd[prefix + ".norm1.gamma"], sd[prefix + ".norm1.beta"])
    z = y.permute(0, 2, 1)
    z = F.linear(z, sd[prefix + ".ff.lin1.weight"], sd[prefix + ".ff.lin1.bias"])
    z = torch.relu(z)
    z = F.linear(z, sd[prefix + ".ff.lin2.weight"], sd[prefix + ".ff.lin2.bias"])
    y2 = y + z.permute(0, 2, 1)
    return layer_norm_channels(y2, sd[prefix + ".norm2.gamma"], sd[prefix + ".norm2.beta"])


def _eqt_decoder(x: torch.Tensor, sd: SD, prefix: str, crops: List[int], tap=None, tap_prefix: str = "") -> torch.Tensor:
    i = 0
    while f"{prefix}.convs.{i}.weight" in sd:
        w = sd[f"{prefix}.convs.{i}.weight"]
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if i in crops:
            x = x[:, :, :-1]
        x = torch.relu(F.conv1d(x, w, sd[f"{prefix}.convs.{i}.bias"], padding=w.shape[2] // 2))
        if tap is not None:
            tap(f"{tap_prefix}_dec{i}", x)
        i += 1
    return x


def eqt_decoder_crops(out_samples: int, n_layers: int) -> List[int]:
    """SeisBench ``Decoder.__init__``: which up-sampling stages drop their last sample."""
    crops = []
    cur = out_samples
    for i in range(n_layers):
        padding = cur % 2
        cur = (cur + padding) // 2
        if padding == 1:
            crops.append(n_layers - 1 - i)
    return crops


def eqtransformer_forward(
    sd: SD, x: torch.Tensor, taps: Optional[Dict[str, torch.Tensor]] = None
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """EQTransformer.forward.  x (B,3,6000) -> (det, P, S) each (B,6000).

    ``taps`` (optional dict) receives named intermediates for layer-by-layer parity checks.
    """
    assert x.ndim == 3 and x.shape[1] == 3
    in_samples = x.shape[2]

    def tap(name, t):
        if taps is not None:
            taps[name] = t.detach().clone()

    with torch.no_grad():
        # Encoder: relu(conv same) -> (odd length: right-pad 1 with -1e10) -> maxpool 2
        i = 0
        while f"encoder.convs.{i}.weight" in sd:
            w = sd[f"encoder.convs.{i}.weight"]
            x = torch.relu(F.conv1d(x, w, sd[f"encoder.convs.{i}.bias"], padding=w.shape[2] // 2))
            if x.shape[2] % 2 == 1:
                x = F.pad(x, (0, 1), "constant", POOL_PAD_VALUE)
            x = F.max_pool1d(x, 2)
            tap(f"enc{i}", x)
            i += 1
        n_enc = i

        # Res-CNN stack
        i = 0
        while f"res_cnn_stack.members.{i}.conv1.weight" in sd:
            p = f"res_cnn_stack.members.{i}"
            ker = sd[p + ".conv1.weight"].shape[2]
            manual = ker != 3  # ker == 2: right-pad one zero, conv padding 0
            pad = 0 if manual else 1
            y = torch.relu(_bn(x, sd, p + ".norm1"))
            if manual:
                y = F.pad(y, (0, 1), "constant", 0)
            y = F.conv1d(y, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=pad)
            y = torch.relu(_bn(y, sd, p + ".norm2"))
            if manual:
                y = F.pad(y, (0, 1), "constant", 0)
            y = F.conv1d(y, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=pad)
            x = x + y
            tap(f"res{i}", x)
            i += 1

        # BiLSTM blocks
        i = 0
        while f"bi_lstm_stack.members.{i}.conv.weight" in sd:
            p = f"bi_lstm_stack.members.{i}"
            x = lstm(x, sd, p + ".lstm", bidirectional=True)
            tap(f"bilstm{i}_lstm", x)
            x = F.conv1d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"])
            x = _bn(x, sd, p + ".norm")
            tap(f"bilstm{i}", x)
            i += 1

        x = transformer(x, sd, "transformer_d0")
        tap("transformer_d0", x)
        x = transformer(x, sd, "transformer_d")
        tap("transformer_d", x)

        crops = eqt_decoder_crops(in_samples, n_enc)
        d = _eqt_decoder(x, sd, "decoder_d", crops, tap, "d0")
        tap("decoder_d", d)
        det = torch.sigmoid(F.conv1d(d, sd["conv_d.weight"], sd["conv_d.bias"], padding=sd["conv_d.weight"].shape[2] // 2))
        outputs = [det.squeeze(1)]

        i = 0
        while f"pick_lstms.{i}.weight_ih_l0" in sd:
            px = lstm(x, sd, f"pick_lstms.{i}", bidirectional=False)
            tap(f"pick{i}_lstm", px)
            px = seq_self_attention(px, sd, f"pick_attentions.{i}", width=3)
            tap(f"pick{i}_attn", px)
            px = _eqt_decoder(px, sd, f"pick_decoders.{i}", crops, tap, f"d{i + 1}")
            w = sd[f"pick_convs.{i}.weight"]
            pred = torch.sigmoid(F.conv1d(px, w, sd[f"pick_convs.{i}.bias"], padding=w.shape[2] // 2))
            outputs.append(pred.squeeze(1))
            i += 1
    return tuple(outputs)


# PhaseNet manual zero pads before the strided down-convolutions (Appendix D #2).
PN_DOWN_PADS = {0: (3, 3), 1: (2, 3), 2: (1, 3), 3: (2, 3)}


def phasenet_forward(sd: SD, x: torch.Tensor, taps: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """PhaseNet.forward.  x (B,3,3001) -> (B,3,3001) softmax, channels P,S,N."""
    assert x.ndim == 3 and x.shape[1] == 3

    def tap(name, t):
        if taps is not None:
            taps[name] = t.detach().clone()

    with torch.no_grad():
        w = sd["inc.weight"]
        x = torch.relu(_bn(F.conv1d(x, w, sd["inc.bias"], padding=w.shape[2] // 2), sd, "in_bn"))
        tap("inc", x)
        skips = []
        depth = 0
        while f"down_branch.{depth}.0.weight" in sd:
            depth += 1
        for i in range(depth):
            w = sd[f"down_branch.{i}.0.weight"]
            x = torch.relu(_bn(F.conv1d(x, w, None, padding=w.shape[2] // 2), sd, f"down_branch.{i}.1"))
            tap(f"down{i}_same", x)
            if f"down_branch.{i}.2.weight" in sd:
                skips.append(x)
                x = F.pad(x, PN_DOWN_PADS[i], "constant", 0)
                x = F.conv1d(x, sd[f"down_branch.{i}.2.weight"], None, stride=4, padding=0)
                x = torch.relu(_bn(x, sd, f"down_branch.{i}.3"))
                tap(f"down{i}_down", x)
        for i, skip in enumerate(skips[::-1]):
            x = F.conv_transpose1d(x, sd[f"up_branch.{i}.0.weight"], None, stride=4)
            x = torch.relu(_bn(x, sd, f"up_branch.{i}.1"))
            x = x[:, :, 1:-2]
            offset = (x.shape[-1] - skip.shape[-1]) // 2
            x = torch.cat([skip, x[:, :, offset : offset + skip.shape[-1]]], dim=1)
            tap(f"up{i}_cat", x)
            w = sd[f"up_branch.{i}.2.weight"]
            x = torch.relu(_bn(F.conv1d(x, w, None, padding=w.shape[2] // 2), sd, f"up_branch.{i}.3"))
            tap(f"up{i}_same", x)
        x = F.conv1d(x, sd["out.weight"], sd["out.bias"])
        return torch.softmax(x, dim=1)


def load_state_dict(path: str) -> SD:
    """Load a SeisBench ``.pt.v1`` pickle (tensors only) and drop the BN counters."""
    sd = torch.load(path, weights_only=True, map_location="cpu")
    return {k: v.float() for k, v in sd.items() if not k.endswith("num_batches_tracked")}


def macs_per_window(sd: SD, kind: str) -> float:
    """Algorithmic MACs per window (conv / convT / linear / LSTM / attention matmuls), Appendix E."""
    if kind == "phasenet":
        L = [3001, 751, 188, 47, 12]
        m = 0.0
        m += sd["inc.weight"].numel() * L[0]
        for i in range(5):
            m += sd[f"down_branch.{i}.0.weight"].numel() * L[i]
            if i < 4:
                m += sd[f"down_branch.{i}.2.weight"].numel() * L[i + 1]
        for i in range(4):
            m += sd[f"up_branch.{i}.0.weight"].numel() * L[4 - i]
            m += sd[f"up_branch.{i}.2.weight"].numel() * L[3 - i]
        m += sd["out.weight"].numel() * L[0]
        return m
    L = 6000
    m = 0.0
    cur = L
    for i in range(7):
        m += sd[f"encoder.convs.{i}.weight"].numel() * cur
        cur = (cur + cur % 2) // 2
    T = cur
    for i in range(7):
        m += 2 * sd[f"res_cnn_stack.members.{i}.conv1.weight"].numel() * T
    for i in range(3):
        p = f"bi_lstm_stack.members.{i}"
        m += 2 * (sd[p + ".lstm.weight_ih_l0"].numel() + sd[p + ".lstm.weight_hh_l0"].numel()) * T
        m += sd[p + ".conv.weight"].numel() * T
    attn = 2 * T * 16 * 32 + T * T * 32 + T * T * 16
    ff = 2 * T * 16 * 128
    m += 2 * (attn + ff)
    lens = [94, 188, 375, 750, 1500, 3000, 6000]
    for pre in ["decoder_d", "pick_decoders.0", "pick_decoders.1"]:
        for i in range(7):
            m += sd[f"{pre}.convs.{i}.weight"].numel() * lens[i]
        m += 88 * L
    for i in range(2):
        m += (sd[f"pick_lstms.{i}.weight_ih_l0"].numel() + sd[f"pick_lstms.{i}.weight_hh_l0"].numel()) * T
        m += 2 * T * 16 * 32 + T * T * 32 + T * 3 * 16
    return m


def state_dict_from_numpy(weights) -> SD:
    """Build the torch state dict from the torch-free VPW1 container (volpick_b200.weights_io.read_vpw)."""
    return {k: torch.from_numpy(v.copy()).float() for k, v in weights.items() if not k.endswith("num_batches_tracked")}
