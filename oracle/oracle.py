"""oracle/oracle.py -- TEST INFRASTRUCTURE (the parity oracle), NOT product code.

A numpy restatement of the reference's CPU algorithm for the hot path
(16 kHz PCM -> log-mel -> FastConformer encoder -> CTC / TDT greedy decode),
written from the reference sources (Frikallo/parakeet.cpp @ 40bbd7e, axiom @
286b305); every function cites the file:line it follows.  It is pinned in
tests/test_oracle.py against

  * the reference's own known-answer tests that apply at this boundary
    (tests/test_all.cpp:759-872 CTCDecode.*, :1003-1030 PositionEmbedding.*,
    :45-129 GroupTimestamps.* / TimestampTypes.*), and
  * golden vectors produced by the UNMODIFIED reference compiled here
    (oracle/_ref/libpkref.so, built by oracle/Makefile; generator script
    tests/golden/make_golden.py), committed under tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.  The product (parakeet.cpp_b200/) never does.

Numerics: everything is float32 like the reference, except where the reference
itself uses double (mel filterbank construction, audio.cpp:40-94; convolution
accumulators, axiom operations.cpp:3074,3252).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------
# Config presets                                     include/parakeet/config.hpp
# ----------------------------------------------------------------------------
@dataclass
class Config:
    """EncoderConfig + heads (config.hpp:9-75); presets at :77-116."""
    mel_bins: int = 80
    sub_channels: int = 256
    d_model: int = 512
    n_layers: int = 17
    n_heads: int = 8
    ff: int = 2048
    conv_k: int = 9
    vocab: int = 1025          # incl. blank = vocab-1
    pred_hidden: int = 640
    lstm_layers: int = 1
    joint_hidden: int = 640
    durations: tuple = (0, 1, 2, 3, 4)
    has_ctc: bool = True
    joint_prefix: str = "tdt_joint_."   # ParakeetTDTCTC (tdt_ctc.cpp:5-9)
    name: str = "tdt-ctc-110m"
    # streaming encoder only (StreamingEncoderConfig, streaming_encoder.hpp:18-24)
    att_context_left: int = 70
    att_context_right: int = 0
    xscaling: bool = False


def make_110m_config() -> Config:          # config.hpp:77-95
    return Config()


def make_tdt_600m_config() -> Config:      # config.hpp:98-116, tdt.cpp:28-32
    return Config(mel_bins=128, d_model=1024, n_layers=24, n_heads=8, ff=4096,
                  vocab=8193, lstm_layers=2, has_ctc=False, joint_prefix="joint_.",
                  name="tdt-600m")


def make_eou_120m_config() -> Config:      # eou.hpp:32-55 (streaming; ParakeetEOU registers "joint_", eou.cpp:9-13)
    return Config(has_ctc=False, joint_prefix="joint_.", att_context_left=70, att_context_right=1, name="eou-120m")


def make_tiny_stream_config() -> Config:
    """Not a reference preset: a small streaming shape for fast unit tests only."""
    return Config(mel_bins=80, sub_channels=64, d_model=128, n_layers=2, n_heads=2, ff=256, vocab=33, pred_hidden=64,
                  joint_hidden=64, has_ctc=False, joint_prefix="joint_.", att_context_left=12, att_context_right=1,
                  name="tiny-stream")


def make_tiny_config() -> Config:
    """Not a reference preset: a small shape for fast unit tests only."""
    return Config(mel_bins=80, sub_channels=64, d_model=128, n_layers=2, n_heads=2, ff=256,
                  vocab=33, pred_hidden=64, joint_hidden=64, name="tiny")


# ----------------------------------------------------------------------------
# Mel front end                                                  src/audio.cpp
# ----------------------------------------------------------------------------
def _hz_to_mel(f: float) -> float:          # audio.cpp:25-30
    return f / (200.0 / 3.0) if f < 1000.0 else 15.0 + math.log(f / 1000.0) / 0.06875177742094912


def _mel_to_hz(m: float) -> float:          # audio.cpp:32-37
    return m * (200.0 / 3.0) if m < 15.0 else 1000.0 * math.exp((m - 15.0) * 0.06875177742094912)


def mel_filterbank(n_freqs: int = 257, n_mels: int = 80, sr: float = 16000.0,
                   f_min: float = 0.0, f_max: float = 8000.0) -> np.ndarray:
    """Slaney filterbank, built in double, stored fp32, (n_freqs, n_mels). audio.cpp:40-94."""
    mel_min, mel_max = _hz_to_mel(f_min), _hz_to_mel(f_max)
    mel_pts = [mel_min + i * (mel_max - mel_min) / (n_mels + 1) for i in range(n_mels + 2)]
    hz = [_mel_to_hz(m) for m in mel_pts]
    freqs = [i * float(sr) / (2.0 * (n_freqs - 1)) for i in range(n_freqs)]
    fb = np.zeros((n_freqs, n_mels), dtype=F32)
    for m in range(n_mels):
        left, center, right = hz[m], hz[m + 1], hz[m + 2]
        enorm = 2.0 / (right - left)
        for f in range(n_freqs):
            fr = freqs[f]
            val = 0.0
            if left <= fr <= center and center > left:
                val = (fr - left) / (center - left)
            elif center < fr <= right and right > center:
                val = (right - fr) / (right - center)
            fb[f, m] = F32(val * enorm)
    return fb


def hann_window(M: int) -> np.ndarray:
    """Symmetric (periodic=False) Hann, double formula cast to fp32. axiom fft.cpp:1117-1142."""
    N = max(M - 1, 1)
    i = np.arange(M, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * math.pi * i / N)).astype(F32)


def n_mel_frames(n_samples: int, hop: int = 160) -> int:
    """center=True STFT: 1 + floor(N / hop) (fft.cpp:1505-1524 with pad n_fft/2 each side)."""
    return 1 + n_samples // hop


def log_mel_unnormalised(pcm: np.ndarray, n_mels: int = 80) -> np.ndarray:
    """Steps 1-5 of preprocess_audio (audio.cpp:100-136) -> (n_mels, n_frames) fp32."""
    n_fft, win_len, hop = 512, 400, 160
    x = np.asarray(pcm, dtype=F32)
    pre = np.empty_like(x)                                   # :104-114
    pre[0] = x[0]
    pre[1:] = x[1:] - F32(0.97) * x[:-1]
    pad = n_fft // 2                                         # fft.cpp:1505-1513 reflect
    sig = np.pad(pre, (pad, pad), mode="reflect")
    win = np.zeros(n_fft, dtype=F32)                         # fft.cpp:1539-1547 centred pad
    lp = (n_fft - win_len) // 2
    win[lp:lp + win_len] = hann_window(win_len)
    n_frames = (len(sig) - n_fft) // hop + 1
    idx = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = sig[idx] * win[None, :]                         # fft.cpp:1561-1575
    spec = np.fft.rfft(frames.astype(F32), n=n_fft, axis=1).astype(np.complex64)
    mag = np.abs(spec).astype(F32)                           # audio.cpp:123-124
    power = (mag * mag).T                                    # (257, n_frames)
    fb = mel_filterbank(n_fft // 2 + 1, n_mels)              # :126-130
    mel = fb.T.astype(F32) @ power                           # :132
    return np.log(mel + F32(5.96046448e-8)).astype(F32)      # :135-136


def preprocess_audio(pcm: np.ndarray, n_mels: int = 80) -> np.ndarray:
    """preprocess_audio (audio.cpp:100-158) -> (n_frames, n_mels) fp32."""
    lm = log_mel_unnormalised(pcm, n_mels)
    n = lm.shape[1]
    mean = lm.mean(axis=1, keepdims=True, dtype=F32)         # :142
    c = lm - mean
    var = (c * c).sum(axis=1, keepdims=True, dtype=F32) / F32(n - 1)   # :146-147 unbiased
    feat = c / (np.sqrt(var) + F32(1e-5))                    # :148 eps outside sqrt
    return np.ascontiguousarray(feat.T.astype(F32))          # :156


def sinc_resample(x, src_rate: int, dst_rate: int):
    """sinc_resample (audio_io.cpp:123-195), in double like the reference: output length ceil(n*up/down); per
    output a 32-tap Kaiser(beta 7.857)-windowed sinc around floor(i * src/dst), cutoff min(1, dst/src), the window
    (not the tap range) widened by src/dst when downsampling, weights renormalised by their sum."""
    x = np.asarray(x, F32)
    if src_rate == dst_rate:
        return x.copy()
    g = math.gcd(src_rate, dst_rate)
    up, down = dst_rate // g, src_rate // g
    n = len(x)
    m = (n * up + down - 1) // down
    HW, BETA = 16, 7.857
    ratio = src_rate / dst_rate
    cutoff = min(1.0, 1.0 / max(ratio, 1.0))
    sample_ratio = dst_rate / src_rate
    width = max(1.0, ratio)

    def i0(v):                                                  # bessel_i0 :101-111 (series with the same stopping rule)
        v = np.asarray(v, np.float64)
        s = np.ones_like(v)
        term = np.ones_like(v)
        done = np.zeros(v.shape, bool)
        for k in range(1, 30):
            term = np.where(done, term, term * (v * v) / (4.0 * k * k))
            s = np.where(done, s, s + term)
            done |= term < 1e-12 * s
        return s

    i = np.arange(m, dtype=np.float64)
    src_pos = i / sample_ratio
    center = np.floor(src_pos).astype(np.int64)
    j = center[:, None] + np.arange(-HW + 1, HW + 1)[None, :]              # taps center-15 .. center+16
    valid = (j >= 0) & (j < n)
    dist = src_pos[:, None] - j
    wpos = dist / width
    valid &= np.abs(wpos) <= HW
    arg = 2.0 * (wpos + HW) / (2.0 * HW) - 1.0                             # kaiser_window :114-121
    val = np.maximum(1.0 - arg * arg, 0.0)
    w = i0(BETA * np.sqrt(val)) / i0(np.float64(BETA))
    xx = dist * cutoff * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(np.abs(xx) < 1e-10, 1.0, np.sin(xx) / xx)
    weight = np.where(valid, sinc * w * cutoff, 0.0)
    xs = x.astype(np.float64)[np.clip(j, 0, max(n - 1, 0))]
    ssum = (xs * weight).sum(axis=1)
    wsum = weight.sum(axis=1)
    return np.where(wsum > 1e-10, ssum / np.where(wsum == 0, 1, wsum), 0.0).astype(F32)


# ----------------------------------------------------------------------------
# axiom primitives                       third_party/axiom/src/tensor/operations.cpp
# ----------------------------------------------------------------------------
def linear(x, w, b=None):
    """nn::Linear: x W^T + b, W = (out, in). axiom linear.cpp:15-27."""
    y = x.astype(F32) @ w.T.astype(F32)
    return y + b if b is not None else y


def layer_norm(x, w, b, eps=1e-5):
    """Biased variance, eps inside sqrt. operations.cpp:1796-1809."""
    mu = x.mean(axis=-1, keepdims=True, dtype=F32)
    c = x - mu
    var = (c * c).mean(axis=-1, keepdims=True, dtype=F32)
    return (c / np.sqrt(var + F32(eps)) * w + b).astype(F32)


def sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x.astype(F32)))).astype(F32)


def silu(x):
    return (x * sigmoid(x)).astype(F32)


def softmax(x, axis=-1):
    """Max-subtracted softmax. cpu_operations.cpp:3101-3231."""
    m = x.max(axis=axis, keepdims=True)
    e = np.exp((x - m).astype(F32))
    return (e / e.sum(axis=axis, keepdims=True, dtype=F32)).astype(F32)


def log_softmax(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    s = (x - m).astype(F32)
    return (s - np.log(np.exp(s).sum(axis=axis, keepdims=True, dtype=F32))).astype(F32)


def conv_out_len(L: int, k: int = 3, s: int = 2, p: int = 1) -> int:
    """operations.cpp:3191-3196: floor((L + 2p - k) / s) + 1."""
    return (L + 2 * p - k) // s + 1


def conv2d(x, w, b, stride, pad, groups=1):
    """Cross-correlation, zero pad, fp64 accumulate. operations.cpp:3133-3326.
    x (C_in, H, W), w (C_out, C_in/g, kh, kw) -> (C_out, H', W')."""
    C_in, H, W = x.shape
    C_out, cg, kh, kw = w.shape
    Ho, Wo = conv_out_len(H, kh, stride, pad), conv_out_len(W, kw, stride, pad)
    xp = np.zeros((C_in, H + 2 * pad, W + 2 * pad), dtype=np.float64)
    xp[:, pad:pad + H, pad:pad + W] = x
    out = np.zeros((C_out, Ho, Wo), dtype=np.float64)
    opg = C_out // groups
    for g in range(groups):
        xs = xp[g * cg:(g + 1) * cg]
        ws = w[g * opg:(g + 1) * opg].astype(np.float64)
        for i in range(kh):
            for j in range(kw):
                patch = xs[:, i:i + stride * Ho:stride, j:j + stride * Wo:stride]   # (cg, Ho, Wo)
                out[g * opg:(g + 1) * opg] += np.einsum("oc,chw->ohw", ws[:, :, i, j], patch)
    if b is not None:
        out += b.astype(np.float64)[:, None, None]
    return out.astype(F32)


def depthwise_conv2d(x, w, b, stride, pad):
    """groups = C specialisation of conv2d (same arithmetic, vectorised)."""
    C, H, W = x.shape
    kh, kw = w.shape[2], w.shape[3]
    Ho, Wo = conv_out_len(H, kh, stride, pad), conv_out_len(W, kw, stride, pad)
    xp = np.zeros((C, H + 2 * pad, W + 2 * pad), dtype=np.float64)
    xp[:, pad:pad + H, pad:pad + W] = x
    out = np.zeros((C, Ho, Wo), dtype=np.float64)
    for i in range(kh):
        for j in range(kw):
            out += w[:, 0, i, j].astype(np.float64)[:, None, None] * \
                xp[:, i:i + stride * Ho:stride, j:j + stride * Wo:stride]
    out += b.astype(np.float64)[:, None, None]
    return out.astype(F32)


def depthwise_conv1d(x, w, b, pad):
    """conv1d groups=C, stride 1. operations.cpp:2960-3127. x (C, T), w (C,1,k)."""
    C, T = x.shape
    k = w.shape[2]
    xp = np.zeros((C, T + 2 * pad), dtype=np.float64)
    xp[:, pad:pad + T] = x
    out = np.zeros((C, T + 2 * pad - k + 1), dtype=np.float64)
    for j in range(k):
        out += w[:, 0, j].astype(np.float64)[:, None] * xp[:, j:j + out.shape[1]]
    out += b.astype(np.float64)[:, None]
    return out.astype(F32)


# ----------------------------------------------------------------------------
# FastConformer encoder                                         src/encoder.cpp
# ----------------------------------------------------------------------------
def sinusoidal_position_embedding(seq_len: int, d_model: int) -> np.ndarray:
    """encoder.cpp:9-30: row r <-> relative position seq_len-1-r; fp32 arithmetic."""
    total = 2 * seq_len - 1
    pe = np.zeros((total, d_model), dtype=F32)
    pos = (F32(seq_len - 1) - np.arange(total, dtype=F32)).astype(F32)
    i = np.arange(0, d_model, 2, dtype=F32)
    div = np.exp(i * (F32(-math.log(F32(10000.0))) / F32(d_model))).astype(F32)
    ang = (pos[:, None] * div[None, :]).astype(F32)
    pe[:, 0::2] = np.sin(ang)
    pe[:, 1::2] = np.cos(ang)[:, : d_model // 2]
    return pe


def conv_subsampling(W, feats, cfg: Config, return_stages=False):
    """ConvSubsampling::forward, encoder.cpp:219-241 (ReLU, not SiLU). feats (T, mel)."""
    p = "encoder_.subsampling_."
    x = feats[None, :, :]                                             # (1, T, F)
    x = np.maximum(conv2d(x, W[p + "conv1_.weight"], W[p + "conv1_.bias"], 2, 1), 0)
    s1 = x
    x = depthwise_conv2d(x, W[p + "dw1_.weight"], W[p + "dw1_.bias"], 2, 1)
    d1 = x
    x = np.maximum(conv2d(x, W[p + "conv2_.weight"], W[p + "conv2_.bias"], 1, 0), 0)
    x = depthwise_conv2d(x, W[p + "dw2_.weight"], W[p + "dw2_.bias"], 2, 1)
    x = np.maximum(conv2d(x, W[p + "conv3_.weight"], W[p + "conv3_.bias"], 1, 0), 0)
    C, T, Fq = x.shape
    flat = np.ascontiguousarray(x.transpose(1, 0, 2)).reshape(T, C * Fq)   # :236-238
    out = linear(flat, W[p + "proj_.weight"], W[p + "proj_.bias"]).astype(F32)
    if return_stages:
        return out, dict(conv1=s1, dw1=d1)
    return out


def feed_forward(W, p, x):
    """FeedForward::forward, encoder.cpp:39-46."""
    h = layer_norm(x, W[p + "norm_.weight"], W[p + "norm_.bias"])
    h = silu(linear(h, W[p + "fc1_.weight"], W[p + "fc1_.bias"]))
    h = linear(h, W[p + "fc2_.weight"], W[p + "fc2_.bias"])
    return (x + h * F32(0.5)).astype(F32)


def rel_shift(x):
    """ConformerAttention::rel_shift, encoder.cpp:85-109, literally. x (H, T, 2T-1)."""
    H, T, P = x.shape
    padded = np.concatenate([np.zeros((H, T, 1), dtype=x.dtype), x], axis=2)   # pad left 1
    padded = padded.reshape(H, P + 1, T)[:, 1:, :]
    return padded.reshape(H, T, P)[:, :, :T]


def conformer_attention(W, p, x, pos_emb, cfg: Config):
    """ConformerAttention::forward / rel_position_attention, encoder.cpp:111-186."""
    T, d = x.shape
    H = cfg.n_heads
    hd = d // H
    h = layer_norm(x, W[p + "norm_.weight"], W[p + "norm_.bias"])
    q = linear(h, W[p + "mha_.q_proj.weight"], W[p + "mha_.q_proj.bias"]).reshape(T, H, hd).transpose(1, 0, 2)
    k = linear(h, W[p + "mha_.k_proj.weight"], W[p + "mha_.k_proj.bias"]).reshape(T, H, hd).transpose(1, 0, 2)
    v = linear(h, W[p + "mha_.v_proj.weight"], W[p + "mha_.v_proj.bias"]).reshape(T, H, hd).transpose(1, 0, 2)
    u = W[p + "pos_bias_u_"].reshape(H, 1, hd)
    vb = W[p + "pos_bias_v_"].reshape(H, 1, hd)
    ac = (q + u) @ k.transpose(0, 2, 1)                                  # (H, T, T)
    pp = linear(pos_emb, W[p + "pos_proj_.weight"]).reshape(2 * T - 1, H, hd).transpose(1, 0, 2)
    bd = rel_shift((q + vb) @ pp.transpose(0, 2, 1))                     # (H, T, T)
    scores = ((ac + bd) * F32(1.0 / math.sqrt(hd))).astype(F32)
    a = softmax(scores, axis=-1)
    o = (a @ v).transpose(1, 0, 2).reshape(T, d)
    o = linear(o, W[p + "mha_.out_proj.weight"], W[p + "mha_.out_proj.bias"])
    return (x + o).astype(F32)


def conformer_conv(W, p, x, cfg: Config):
    """ConformerConvModule::forward, encoder.cpp:59-75; BN eval axiom normalization.cpp:48-104."""
    d = x.shape[1]
    h = layer_norm(x, W[p + "norm_.weight"], W[p + "norm_.bias"])
    h = linear(h, W[p + "pointwise_conv1_.weight"][:, :, 0], W[p + "pointwise_conv1_.bias"])  # (T, 2d)
    h = (h[:, :d] * sigmoid(h[:, d:])).astype(F32)                       # GLU operations.cpp:1450-1476
    h = depthwise_conv1d(h.T, W[p + "depthwise_conv_.weight"], W[p + "depthwise_conv_.bias"],
                         (cfg.conv_k - 1) // 2).T                        # (T, d)
    mean, var = W[p + "batch_norm_.running_mean"], W[p + "batch_norm_.running_var"]
    h = ((h - mean) / np.sqrt(var + F32(1e-5)) * W[p + "batch_norm_.weight"] + W[p + "batch_norm_.bias"]).astype(F32)
    h = silu(h)
    h = linear(h, W[p + "pointwise_conv2_.weight"][:, :, 0], W[p + "pointwise_conv2_.bias"])
    return (x + h).astype(F32)


def conformer_block(W, i, x, pos_emb, cfg: Config):
    """ConformerBlock::forward, encoder.cpp:196-204."""
    p = f"encoder_.layers_.{i}."
    x = feed_forward(W, p + "ffn1_.", x)
    x = conformer_attention(W, p + "attn_.", x, pos_emb, cfg)
    x = conformer_conv(W, p + "conv_.", x, cfg)
    x = feed_forward(W, p + "ffn2_.", x)
    return layer_norm(x, W[p + "final_norm_.weight"], W[p + "final_norm_.bias"])


def encoder_forward(W, feats, cfg: Config, return_layers=False):
    """FastConformerEncoder::forward, encoder.cpp:253-271. feats (n_frames, mel) -> (T', d)."""
    x = conv_subsampling(W, feats, cfg)
    pos = sinusoidal_position_embedding(x.shape[0], x.shape[1])
    layers = []
    sub = x
    for i in range(cfg.n_layers):
        x = conformer_block(W, i, x, pos, cfg)
        if return_layers:
            layers.append(x)
    if return_layers:
        return x, sub, np.stack(layers)
    return x


def encoder_len(n_frames: int) -> int:
    """Three stride-2 k3 p1 stages (encoder.cpp:208-217)."""
    t = n_frames
    for _ in range(3):
        t = conv_out_len(t)
    return t


# ----------------------------------------------------------------------------
# CTC head + greedy                                               src/ctc.cpp
# ----------------------------------------------------------------------------
def ctc_log_probs(W, enc):
    """CTCDecoder::forward, ctc.cpp:12-25 (k=1 Conv1d + log_softmax)."""
    w = W["ctc_decoder_.proj_.weight"][:, :, 0]
    return log_softmax(linear(enc, w, W["ctc_decoder_.proj_.bias"]), axis=-1)


def first_argmax(row) -> int:
    """Strict '>' scan: first maximum wins. ctc.cpp:59-66."""
    return int(np.argmax(row))     # np.argmax returns the first occurrence


def ctc_greedy_decode(lp, blank=1024):
    """ctc_greedy_decode, ctc.cpp:40-75. lp (T, V) -> list[int]."""
    out, prev = [], -1
    for t in range(lp.shape[0]):
        best = first_argmax(lp[t])
        if best != blank and best != prev:
            out.append(best)
        prev = best
    return out


def ctc_greedy_decode_with_timestamps(lp, blank=1024):
    """ctc.cpp:79-127 -> list of (id, start_frame, end_frame, confidence)."""
    toks, prev = [], -1
    T = lp.shape[0]
    for t in range(T):
        best = first_argmax(lp[t])
        if best != prev:
            if prev != -1 and prev != blank and toks:
                toks[-1][2] = t - 1
            if best != blank:
                toks.append([best, t, t, float(np.exp(F32(lp[t, best])))])
        prev = best
    if toks:
        toks[-1][2] = T - 1
    return [tuple(t) for t in toks]


# ----------------------------------------------------------------------------
# Prediction net + TDT joint + greedy      src/rnnt.cpp src/lstm.cpp src/tdt.cpp
# ----------------------------------------------------------------------------
def lstm_cell(W, p, x, h, c):
    """LSTMCell::forward, lstm.cpp:11-29: gates i,f,g,o; one merged bias."""
    g = linear(x, W[p + "input_proj_.weight"], W[p + "input_proj_.bias"]) + linear(h, W[p + "hidden_proj_.weight"])
    H = h.shape[-1]
    i, f, gg, o = g[..., :H], g[..., H:2 * H], g[..., 2 * H:3 * H], g[..., 3 * H:]
    c2 = (sigmoid(f) * c + sigmoid(i) * np.tanh(gg)).astype(F32)
    h2 = (sigmoid(o) * np.tanh(c2)).astype(F32)
    return h2, c2


def prediction_step(W, token, states, cfg: Config):
    """RNNTPrediction::step rnnt.cpp:22-28 -> LSTM::step lstm.cpp:40-49."""
    x = W["prediction_.embed_.weight"][token]
    new = []
    for l in range(cfg.lstm_layers):
        h, c = lstm_cell(W, f"prediction_.lstm_.cells_.{l}.", x, *states[l])
        new.append((h, c))
        x = h
    return x, new


def tdt_joint(W, enc_t, pred, cfg: Config):
    """TDTJoint::forward, tdt.cpp:15-24 (pred_proj_ has no bias)."""
    p = cfg.joint_prefix
    z = linear(enc_t, W[p + "enc_proj_.weight"], W[p + "enc_proj_.bias"]) + linear(pred, W[p + "pred_proj_.weight"])
    z = np.maximum(z, 0).astype(F32)
    lab = log_softmax(linear(z, W[p + "label_proj_.weight"], W[p + "label_proj_.bias"]))
    dur = log_softmax(linear(z, W[p + "duration_proj_.weight"], W[p + "duration_proj_.bias"]))
    return lab, dur


def tdt_greedy_decode(W, enc, cfg: Config, max_symbols=10, with_timestamps=False, max_steps=100000):
    """tdt_greedy_decode(_with_timestamps), tdt.cpp:36-110 / :122-201, quirks and all:
    state revert on blank, blank advances max(skip,1), no forced advance after
    max_symbols zero-duration emissions (the frame is re-entered)."""
    T = enc.shape[0]
    blank = cfg.vocab - 1
    H = cfg.pred_hidden
    states = [(np.zeros(H, F32), np.zeros(H, F32)) for _ in range(cfg.lstm_layers)]
    token, t, out, steps = blank, 0, [], 0
    while t < T:
        for _sym in range(max_symbols):
            steps += 1
            if steps > max_steps:
                raise RuntimeError("tdt_greedy_decode: livelock (reference quirk iv)")
            saved = states
            pred, states = prediction_step(W, token, states, cfg)
            lab, dur = tdt_joint(W, enc[t], pred, cfg)
            tok = first_argmax(lab)
            di = first_argmax(dur)
            skip = cfg.durations[di] if di < len(cfg.durations) else 1
            if tok == blank:
                states = saved
                t += max(skip, 1)
                break
            if with_timestamps:
                out.append((tok, t, min(t + max(skip, 1) - 1, T - 1), float(np.exp(F32(lab[tok])))))
            else:
                out.append(tok)
            token = tok
            if skip > 0:
                t += skip
                break
    return out


# ----------------------------------------------------------------------------
# Streaming path (eou-120m)         src/audio.cpp:170-259, src/streaming_encoder.cpp, src/eou.cpp
# ----------------------------------------------------------------------------
class StreamingPreprocessor:
    """StreamingAudioPreprocessor (audio.cpp:172-259): carried pre-emphasis sample, overlap buffer of the
    samples past the last complete frame, center=False STFT (no reflect padding), log-mel WITHOUT
    normalisation (the running statistics of the header are never used)."""

    def __init__(self, n_mels: int = 80):
        self.n_mels = n_mels
        self.last = F32(0.0)
        self.overlap = np.zeros(0, F32)

    def process_chunk(self, samples):
        n_fft, win_len, hop = 512, 400, 160
        x = np.asarray(samples, F32)
        pre = np.empty_like(x)                                    # :206-213
        prev = self.last
        if len(x):
            pre[0] = x[0] - F32(0.97) * prev
            pre[1:] = x[1:] - F32(0.97) * x[:-1]
            self.last = x[-1]
        buf = np.concatenate([self.overlap, pre]).astype(F32)      # :216-221
        total = len(buf)
        if total < win_len:                                        # :225-229
            self.overlap = buf
            return None
        n_frames = (total - win_len) // hop + 1                    # :231-236
        consumed = (n_frames - 1) * hop + win_len                  # :239-240
        self.overlap = buf[consumed:].copy()
        sig = buf[:consumed]
        # fft::stft(center=False) frames are n_fft = 512 long while `consumed` was sized with win_length = 400,
        # so the STFT returns one frame fewer than n_frames (13-14 per 2560-sample chunk) and the overlap
        # buffer holds only the samples past `consumed` -- both quirks kept.
        frames = stft_frames_no_center(sig, n_fft, win_len, hop)
        spec = np.fft.rfft(frames.astype(F32), n=n_fft, axis=1).astype(np.complex64)
        mag = np.abs(spec).astype(F32)
        power = (mag * mag).T
        fb = mel_filterbank(n_fft // 2 + 1, self.n_mels)
        mel = fb.T.astype(F32) @ power
        return np.ascontiguousarray(np.log(mel + F32(5.96046448e-8)).astype(F32).T)   # (n_frames, n_mels)


def stft_frames_no_center(sig, n_fft, win_len, hop):
    """Frames of fft::stft(center=False) (fft.cpp:1478-1605): frame count (len - n_fft)//hop + 1 over the
    signal as given, each frame n_fft samples times the centre-padded window."""
    win = np.zeros(n_fft, dtype=F32)
    lp = (n_fft - win_len) // 2
    win[lp:lp + win_len] = hann_window(win_len)
    if len(sig) < n_fft:                       # fft.cpp:1516-1521: the reference throws here
        raise ValueError("stft: signal length (%d) is less than n_fft (%d)" % (len(sig), n_fft))
    n_frames = (len(sig) - n_fft) // hop + 1
    idx = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    return sig[idx] * win[None, :]


class StreamEncoderCache:
    """EncoderCache / BlockCache (streaming_encoder.hpp:28-44)."""

    def __init__(self, n_layers):
        self.sub = None                                   # leftover mel frames (n, mel)
        self.conv = [None] * n_layers                     # (d, k-1) GLU outputs
        self.k = [None] * n_layers                        # (H, n, hd)
        self.v = [None] * n_layers
        self.frames_seen = 0


def stream_subsample_cached(W, feats, cache: StreamEncoderCache, cfg: Config):
    """CausalConvSubsampling::forward_cached (streaming_encoder.cpp:339-378): prepend the leftover frames,
    run the ordinary (zero-padded) subsampling on the largest multiple of 8 frames, keep the rest."""
    mel_in = feats if cache.sub is None else np.concatenate([cache.sub, feats], axis=0)
    total = mel_in.shape[0]
    consumable = (total // 8) * 8
    if consumable == 0:
        cache.sub = np.ascontiguousarray(mel_in)
        return None
    cache.sub = np.ascontiguousarray(mel_in[consumable:]) if total > consumable else None
    return conv_subsampling(W, mel_in[:consumable], cfg)          # ReLU variant (eou preset default)


def stream_attention_cached(W, p, x, pos_emb, cache: StreamEncoderCache, li, cfg: Config, apply_context_mask=False):
    """StreamingConformerAttention::forward_cached (streaming_encoder.cpp:160-272): K/V of the chunk appended to
    the cache (trimmed to att_context_left AFTER use as this chunk's keys), position scores WITHOUT rel_shift
    (right-most kv_len columns), then the bounded-context mask -- which has NO EFFECT on the reference's CPU
    path: the mask is built as a float32 tensor of 1.0 / 0.0 (:246-258) and axiom's CPU masked_fill reads the
    first BYTE of every mask element (cpu_operations.cpp:2699-2701, :2719-2721); the low byte of 1.0f
    (0x3F800000, little endian) is 0, so nothing is ever filled.  The compiled reference therefore attends to
    every cached key and to the whole chunk.  apply_context_mask=True gives the intended behaviour."""
    C, d = x.shape
    H = cfg.n_heads
    hd = d // H
    h = layer_norm(x, W[p + "norm_.weight"], W[p + "norm_.bias"])
    q = linear(h, W[p + "mha_.q_proj.weight"], W[p + "mha_.q_proj.bias"]).reshape(C, H, hd).transpose(1, 0, 2)
    k = linear(h, W[p + "mha_.k_proj.weight"], W[p + "mha_.k_proj.bias"]).reshape(C, H, hd).transpose(1, 0, 2)
    v = linear(h, W[p + "mha_.v_proj.weight"], W[p + "mha_.v_proj.bias"]).reshape(C, H, hd).transpose(1, 0, 2)
    if cache.k[li] is not None:                                           # :185-189
        k = np.concatenate([cache.k[li], k], axis=1)
        v = np.concatenate([cache.v[li], v], axis=1)
    kv = k.shape[1]
    L, Rr = cfg.att_context_left, cfg.att_context_right
    if kv > L:                                                            # :193-208
        cache.k[li], cache.v[li] = k[:, kv - L:].copy(), v[:, kv - L:].copy()
    else:
        cache.k[li], cache.v[li] = k.copy(), v.copy()
    u = W[p + "pos_bias_u_"].reshape(H, 1, hd)
    vb = W[p + "pos_bias_v_"].reshape(H, 1, hd)
    ac = (q + u) @ k.transpose(0, 2, 1)                                   # (H, C, kv)
    P = pos_emb.shape[0]
    pp = linear(pos_emb, W[p + "pos_proj_.weight"]).reshape(P, H, hd).transpose(1, 0, 2)
    ps = (q + vb) @ pp.transpose(0, 2, 1)                                 # (H, C, P), no rel_shift
    if P > kv:                                                            # :224-232
        ps = ps[:, :, P - kv:]
    scores = ((ac + ps) * F32(1.0 / math.sqrt(hd))).astype(F32)
    if apply_context_mask and (L >= 0 or Rr >= 0):                        # :239-262 (see docstring)
        qi = np.arange(C)[:, None] + (kv - C)
        dist = qi - np.arange(kv)[None, :]
        masked = (dist > L) | (-dist > Rr)
        scores = np.where(masked[None, :, :], F32(-1e9), scores).astype(F32)
    a = softmax(scores, axis=-1)
    o = (a @ v).transpose(1, 0, 2).reshape(C, d)
    o = linear(o, W[p + "mha_.out_proj.weight"], W[p + "mha_.out_proj.bias"])
    return (x + o).astype(F32)


def stream_conv_cached(W, p, x, cache: StreamEncoderCache, li, cfg: Config):
    """CausalConformerConvModule::forward_cached (streaming_encoder.cpp:41-80): the k-1 previous GLU outputs
    (zeros for the first chunk) are prepended, the depthwise conv runs without padding."""
    d = x.shape[1]
    h = layer_norm(x, W[p + "norm_.weight"], W[p + "norm_.bias"])
    h = linear(h, W[p + "pointwise_conv1_.weight"][:, :, 0], W[p + "pointwise_conv1_.bias"])
    h = (h[:, :d] * sigmoid(h[:, d:])).astype(F32).T                      # (d, C)
    cl = cfg.conv_k - 1
    left = cache.conv[li] if cache.conv[li] is not None else np.zeros((d, cl), F32)
    h = np.concatenate([left, h], axis=1)
    cache.conv[li] = np.ascontiguousarray(h[:, h.shape[1] - cl:])
    h = depthwise_conv1d(h, W[p + "depthwise_conv_.weight"], W[p + "depthwise_conv_.bias"], 0).T   # (C, d)
    mean, var = W[p + "batch_norm_.running_mean"], W[p + "batch_norm_.running_var"]
    h = ((h - mean) / np.sqrt(var + F32(1e-5)) * W[p + "batch_norm_.weight"] + W[p + "batch_norm_.bias"]).astype(F32)
    h = silu(h)
    h = linear(h, W[p + "pointwise_conv2_.weight"][:, :, 0], W[p + "pointwise_conv2_.bias"])
    return (x + h).astype(F32)


def stream_encoder_chunk(W, feats, cache: StreamEncoderCache, cfg: Config):
    """StreamingFastConformerEncoder::forward_chunk (streaming_encoder.cpp:425-472). feats (n, mel) -> (C, d) | None."""
    x = stream_subsample_cached(W, feats, cache, cfg)
    if x is None:
        return None
    if cfg.xscaling:
        x = (x * F32(math.sqrt(cfg.d_model))).astype(F32)
    C = x.shape[0]
    pos = sinusoidal_position_embedding(cfg.att_context_left + C, cfg.d_model)      # :451-452
    for i in range(cfg.n_layers):
        p = f"encoder_.layers_.{i}."
        x = feed_forward(W, p + "ffn1_.", x)
        x = stream_attention_cached(W, p + "attn_.", x, pos, cache, i, cfg)
        x = stream_conv_cached(W, p + "conv_.", x, cache, i, cfg)
        x = feed_forward(W, p + "ffn2_.", x)
        x = layer_norm(x, W[p + "final_norm_.weight"], W[p + "final_norm_.bias"])
    cache.frames_seen += C
    return x


class StreamDecodeState:
    """StreamingDecodeState (eou.hpp:80-87)."""

    def __init__(self, cfg: Config):
        H = cfg.pred_hidden
        self.states = [(np.zeros(H, F32), np.zeros(H, F32)) for _ in range(cfg.lstm_layers)]
        self.token = cfg.vocab - 1
        self.frame_offset = 0
        self.tokens = []


def stream_decode_chunk(W, enc, st: StreamDecodeState, cfg: Config, max_symbols=10, max_steps=100000):
    """rnnt_streaming_decode_chunk (eou.cpp:17-98): the offline TDT loop per chunk with carried LSTM state and
    absolute frame numbers; end frame is NOT clamped to the chunk; same no-forced-advance quirk."""
    C = enc.shape[0]
    blank = cfg.vocab - 1
    new, t, steps = [], 0, 0
    base = st.frame_offset
    while t < C:
        for _sym in range(max_symbols):
            steps += 1
            if steps > max_steps:
                raise RuntimeError("stream_decode_chunk: livelock")
            saved = st.states
            pred, st.states = prediction_step(W, st.token, st.states, cfg)
            lab, dur = tdt_joint(W, enc[t], pred, cfg)
            tok = first_argmax(lab)
            di = first_argmax(dur)
            skip = cfg.durations[di] if di < len(cfg.durations) else 1
            if tok == blank:
                st.states = saved
                t += max(skip, 1)
                break
            new.append((tok, base + t, base + t + max(skip, 1) - 1, float(np.exp(F32(lab[tok])))))
            st.tokens.append(tok)
            st.token = tok
            if skip > 0:
                t += skip
                break
    st.frame_offset += C
    return new


# ----------------------------------------------------------------------------
# Phrase-boosted decode                                    src/phrase_boost.cpp
# ----------------------------------------------------------------------------
class ContextTrie:
    """ContextTrie (phrase_boost.cpp:9-66): node 0 = root; active-state sets always contain the root."""

    def __init__(self, phrases=()):
        self.children = [{}]
        self.is_end = [False]
        for ids in phrases:
            self.insert(ids)

    def insert(self, ids):                                     # :11-27
        if not len(ids):
            return
        node = 0
        for t in ids:
            t = int(t)
            nxt = self.children[node].get(t)
            if nxt is None:
                nxt = len(self.children)
                self.children[node][t] = nxt
                self.children.append({})
                self.is_end.append(False)
            node = nxt
        self.is_end[node] = True

    def boosted(self, active):                                 # get_boosted_tokens :40-51
        out = set()
        for st in active:
            if 0 <= st < len(self.children):
                out.update(self.children[st].keys())
        return out

    def advance(self, active, token):                          # :53-66
        nxt = {0}
        for st in active:
            if 0 <= st < len(self.children) and token in self.children[st]:
                nxt.add(self.children[st][token])
        return nxt


def _boosted_argmax(row, boosted, boost):
    """argmax_v(row[v] + boost * [v in boosted]), first maximum (strict '>' scan, phrase_boost.cpp:94-102)."""
    v = np.asarray(row, F32).copy()
    if boosted:
        idx = np.fromiter(boosted, dtype=np.int64)
        idx = idx[idx < len(v)]
        v[idx] = v[idx] + F32(boost)
    return int(np.argmax(v))


def ctc_greedy_decode_with_timestamps_boosted(lp, trie: ContextTrie, boost=5.0, blank=1024):
    """ctc_greedy_decode_with_timestamps_boosted (phrase_boost.cpp:122-176): the plain CTC loop with the boosted
    argmax; the trie advances on every emission; confidence = exp of the UNboosted log-prob."""
    T = lp.shape[0]
    out, prev, active = [], -1, {0}
    for t in range(T):
        best = _boosted_argmax(lp[t], trie.boosted(active), boost)
        if best != prev:
            if prev != -1 and prev != blank and out:
                out[-1] = (out[-1][0], out[-1][1], t - 1, out[-1][3])
            if best != blank:
                out.append((best, t, t, float(np.exp(F32(lp[t, best])))))
                active = trie.advance(active, best)
        prev = best
    if out:
        out[-1] = (out[-1][0], out[-1][1], T - 1, out[-1][3])
    return out


def tdt_greedy_decode_with_timestamps_boosted(W, enc, cfg: Config, trie: ContextTrie, boost=5.0, max_symbols=10,
                                              max_steps=100000):
    """tdt_greedy_decode_with_timestamps_boosted (phrase_boost.cpp:266-352): tdt_greedy_decode with the boosted
    label argmax (durations are not boosted), raw log-prob confidence, end frame clamped to T-1."""
    T = enc.shape[0]
    blank = cfg.vocab - 1
    H = cfg.pred_hidden
    states = [(np.zeros(H, F32), np.zeros(H, F32)) for _ in range(cfg.lstm_layers)]
    token, t, out, steps, active = blank, 0, [], 0, {0}
    while t < T:
        for _sym in range(max_symbols):
            steps += 1
            if steps > max_steps:
                raise RuntimeError("tdt_greedy_decode_boosted: livelock")
            saved = states
            pred, states = prediction_step(W, token, states, cfg)
            lab, dur = tdt_joint(W, enc[t], pred, cfg)
            tok = _boosted_argmax(lab, trie.boosted(active), boost)
            di = first_argmax(dur)
            skip = cfg.durations[di] if di < len(cfg.durations) else 1
            if tok == blank:
                states = saved
                t += max(skip, 1)
                break
            out.append((tok, t, min(t + max(skip, 1) - 1, T - 1), float(np.exp(F32(lab[tok])))))
            active = trie.advance(active, tok)
            token = tok
            if skip > 0:
                t += skip
                break
    return out


def tokenizer_encode(text, pieces):
    """Tokenizer::encode (vocab.cpp:76-117): prepend U+2581, spaces -> U+2581, greedy longest match over the
    pieces on BYTES, unknown bytes skipped."""
    if not pieces or not text:
        return []
    table = {}
    for i, pc in enumerate(pieces):
        table[pc.encode("utf-8")] = i                # later duplicates overwrite, like operator[] (vocab.cpp:70)
    max_len = max(len(k) for k in table)
    data = (SP_MARK + text.replace(" ", SP_MARK)).encode("utf-8")
    out, pos = [], 0
    while pos < len(data):
        for ln in range(min(max_len, len(data) - pos), 0, -1):
            tid = table.get(data[pos:pos + ln])
            if tid is not None:
                out.append(tid)
                pos += ln
                break
        else:
            pos += 1
    return out


# ----------------------------------------------------------------------------
# Tokenizer + word timestamps (host-side)    src/vocab.cpp src/timestamp.cpp
# ----------------------------------------------------------------------------
SP_MARK = "▁"


def load_vocab(path):
    """Tokenizer::load, vocab.cpp:10-27 (piece<TAB>score lines)."""
    pieces = []
    with open(path, "rb") as f:
        for raw in f.read().split(b"\n"):
            line = raw.decode("utf-8", errors="surrogateescape")
            if "\t" in line:
                pieces.append(line.split("\t")[0])
            elif line:
                pieces.append(line)
    return pieces


def detokenize(ids, pieces):
    """Tokenizer::decode, vocab.cpp:29-64."""
    s = "".join(pieces[i] if 0 <= i < len(pieces) else f"[{i}]" for i in ids)
    s = s.replace(SP_MARK, " ")
    return s[1:] if s.startswith(" ") else s


def group_timestamps(tokens, pieces):
    """group_timestamps(Words), timestamp.cpp:24-75; frame = 0.08 s (timestamp.hpp:31-35).
    tokens: iterable of (id, start, end, conf) -> list of (word, start_s, end_s, conf)."""
    tokens = list(tokens)
    if not tokens:
        return []
    words, cur = [], ""
    ws, we, wc = tokens[0][1], tokens[0][2], 1.0
    f2s = lambda fr: float(F32(fr) * F32(0.08))
    for (tid, st, en, cf) in tokens:
        if tid < 0 or tid >= len(pieces):
            continue
        piece = pieces[tid]
        new_word = piece.startswith(SP_MARK)
        if new_word and cur:
            words.append((cur, f2s(ws), f2s(we), wc))
            cur, ws, wc = "", st, 1.0
        cur += piece[1:] if new_word else piece
        we = en
        wc = min(wc, cf)
    if cur:
        words.append((cur, f2s(ws), f2s(we), wc))
    return words


# ----------------------------------------------------------------------------
# Whole path (Transcriber::transcribe, transcribe.hpp:99-179), one utterance
# ----------------------------------------------------------------------------
def transcribe(W, pcm, cfg: Config, decoder="ctc", timestamps=False):
    feats = preprocess_audio(pcm, cfg.mel_bins)
    enc = encoder_forward(W, feats, cfg)
    if decoder == "ctc":
        lp = ctc_log_probs(W, enc)
        return ctc_greedy_decode_with_timestamps(lp, cfg.vocab - 1) if timestamps \
            else ctc_greedy_decode(lp, cfg.vocab - 1)
    return tdt_greedy_decode(W, enc, cfg, with_timestamps=timestamps)
