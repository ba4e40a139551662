"""oracle/refbind.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libpkref.so,
the UNMODIFIED reference compiled from /root/reference by oracle/Makefile plus the
C-ABI harness oracle/ref_harness.cpp.  Importers: tests/, tests/golden/make_golden.py,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpkref.so")
_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_d = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.pkref_last_error.restype = C.c_char_p
        L.pkref_load.restype = C.c_void_p
        L.pkref_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.pkref_load_custom.restype = C.c_void_p
        L.pkref_load_custom.argtypes = [C.c_char_p, C.c_char_p] + [C.c_int] * 10
        L.pkref_free.argtypes = [C.c_void_p]
        L.pkref_mel.argtypes = [_f, C.c_int64, C.c_int, _f]
        L.pkref_posemb.argtypes = [C.c_int, C.c_int, _f]
        L.pkref_encode.argtypes = [C.c_void_p, _f, C.c_int, C.c_int, _f]
        L.pkref_encode_layers.argtypes = [C.c_void_p, _f, C.c_int, C.c_int, _f, _f]
        L.pkref_ctc_logprobs.argtypes = [C.c_void_p, _f, C.c_int, C.c_int, _f]
        L.pkref_ctc_greedy.argtypes = [_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i, _i, _i, _f, _i]
        L.pkref_tdt_greedy.argtypes = [C.c_void_p, _f, C.c_int, C.c_int, C.c_int, C.c_int, _i, _i, _i, _f]
        L.pkref_detok.argtypes = [C.c_void_p, _i, C.c_int, C.c_char_p, C.c_int]
        L.pkref_group_words.argtypes = [C.c_void_p, _i, _i, _i, _f, C.c_int, C.c_char_p, C.c_int, _f, _f, _f]
        L.pkref_transcribe.argtypes = [C.c_void_p, _f, C.c_int64, C.c_int, C.c_int, _i, _d]
        _lib = L
    return _lib


def _check(rc):
    if rc is None or rc < 0:
        raise RuntimeError("reference: " + lib().pkref_last_error().decode())
    return rc


def mel(pcm, n_mels=80):
    pcm = np.ascontiguousarray(pcm, np.float32)
    out = np.zeros((1 + len(pcm) // 160 + 2, n_mels), np.float32)
    n = _check(lib().pkref_mel(pcm, len(pcm), n_mels, out))
    return out[:n].copy()


def posemb(T, d):
    out = np.zeros((2 * T - 1, d), np.float32)
    _check(lib().pkref_posemb(T, d, out))
    return out


def ctc_greedy(lp, blank, with_ts=False):
    lp = np.ascontiguousarray(lp, np.float32)
    if lp.ndim == 2:
        lp = lp[None]
    B, T, V = lp.shape
    ids = np.zeros((B, T), np.int32); st = np.zeros((B, T), np.int32); en = np.zeros((B, T), np.int32)
    cf = np.zeros((B, T), np.float32); lens = np.zeros(B, np.int32)
    _check(lib().pkref_ctc_greedy(lp, B, T, V, blank, int(with_ts), ids, st, en, cf, lens))
    res = []
    for b in range(B):
        n = lens[b]
        if with_ts:
            res.append([(int(ids[b, i]), int(st[b, i]), int(en[b, i]), float(cf[b, i])) for i in range(n)])
        else:
            res.append([int(x) for x in ids[b, :n]])
    return res


class RefModel:
    """Reference ParakeetTDTCTC (preset 0) / ParakeetTDT (preset 1) + Tokenizer."""

    def __init__(self, weights_path, vocab_path="", preset=0, cfg=None):
        if cfg is not None:   # explicit dimensions (oracle.Config), ParakeetTDTCTC layout
            self.h = lib().pkref_load_custom(weights_path.encode(), (vocab_path or "").encode(), cfg.mel_bins,
                                             cfg.sub_channels, cfg.d_model, cfg.n_layers, cfg.n_heads, cfg.ff,
                                             cfg.vocab, cfg.pred_hidden, cfg.lstm_layers, cfg.joint_hidden)
        else:
            self.h = lib().pkref_load(weights_path.encode(), (vocab_path or "").encode(), preset)
        if not self.h:
            raise RuntimeError("reference: " + lib().pkref_last_error().decode())

    def close(self):
        if self.h:
            lib().pkref_free(self.h)
            self.h = None

    def encode(self, feats, d_model):
        feats = np.ascontiguousarray(feats, np.float32)
        out = np.zeros((feats.shape[0], d_model), np.float32)
        T = _check(lib().pkref_encode(self.h, feats, feats.shape[0], feats.shape[1], out))
        return out[:T].copy()

    def encode_layers(self, feats, d_model, n_layers, T):
        feats = np.ascontiguousarray(feats, np.float32)
        sub = np.zeros((T, d_model), np.float32)
        lay = np.zeros((n_layers, T, d_model), np.float32)
        _check(lib().pkref_encode_layers(self.h, feats, feats.shape[0], feats.shape[1], sub, lay))
        return sub, lay

    def ctc_logprobs(self, enc, vocab):
        enc = np.ascontiguousarray(enc, np.float32)
        out = np.zeros((enc.shape[0], vocab), np.float32)
        _check(lib().pkref_ctc_logprobs(self.h, enc, enc.shape[0], enc.shape[1], out))
        return out

    def tdt_greedy(self, enc, with_ts=False, cap=8192):
        enc = np.ascontiguousarray(enc, np.float32)
        ids = np.zeros(cap, np.int32); st = np.zeros(cap, np.int32); en = np.zeros(cap, np.int32)
        cf = np.zeros(cap, np.float32)
        n = _check(lib().pkref_tdt_greedy(self.h, enc, enc.shape[0], enc.shape[1], int(with_ts), cap, ids, st, en, cf))
        if with_ts:
            return [(int(ids[i]), int(st[i]), int(en[i]), float(cf[i])) for i in range(n)]
        return [int(x) for x in ids[:n]]

    def detok(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        buf = C.create_string_buffer(1 << 16)
        _check(lib().pkref_detok(self.h, ids, len(ids), buf, len(buf)))
        return buf.value.decode("utf-8")

    def group_words(self, toks):
        n = len(toks)
        ids = np.array([t[0] for t in toks], np.int32); st = np.array([t[1] for t in toks], np.int32)
        en = np.array([t[2] for t in toks], np.int32); cf = np.array([t[3] for t in toks], np.float32)
        buf = C.create_string_buffer(1 << 16)
        ws = np.zeros(max(n, 1), np.float32); we = np.zeros(max(n, 1), np.float32); wc = np.zeros(max(n, 1), np.float32)
        k = _check(lib().pkref_group_words(self.h, ids, st, en, cf, n, buf, len(buf), ws, we, wc))
        words = buf.value.decode("utf-8").split("\n")[:k]
        return [(words[i], float(ws[i]), float(we[i]), float(wc[i])) for i in range(k)]

    def transcribe(self, pcm, decoder="ctc", cap=8192):
        """Reference Transcriber::transcribe path, stage-timed -> (ids, [pre_ms, enc_ms, dec_ms])."""
        pcm = np.ascontiguousarray(pcm, np.float32)
        ids = np.zeros(cap, np.int32); ms = np.zeros(3, np.float64)
        n = _check(lib().pkref_transcribe(self.h, pcm, len(pcm), 0 if decoder == "ctc" else 1, cap, ids, ms))
        return [int(x) for x in ids[:n]], ms.tolist()


class RefStream:
    """One stream through the compiled reference's STREAMING path (oracle/ref_harness_stream.cpp):
    StreamingAudioPreprocessor -> StreamingFastConformerEncoder::forward_chunk -> rnnt_streaming_decode_chunk."""

    def __init__(self, weights_path, cfg):
        L = lib()
        L.pkref_stream_new.restype = C.c_void_p
        L.pkref_stream_new.argtypes = [C.c_char_p] + [C.c_int] * 12
        L.pkref_stream_last_error.restype = C.c_char_p
        L.pkref_stream_free.argtypes = [C.c_void_p]
        L.pkref_stream_chunk.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        self.L, self.cfg = L, cfg
        self.h = L.pkref_stream_new(weights_path.encode(), cfg.mel_bins, cfg.sub_channels, cfg.d_model, cfg.n_layers, cfg.n_heads,
                                    cfg.ff, cfg.vocab, cfg.pred_hidden, cfg.lstm_layers, cfg.joint_hidden, cfg.att_context_left,
                                    cfg.att_context_right)
        if not self.h:
            raise RuntimeError("pkref_stream_new: " + L.pkref_stream_last_error().decode())

    def close(self):
        if self.h:
            self.L.pkref_stream_free(self.h)
            self.h = None

    def chunk(self, pcm):
        """-> (feats (n, mel) | None, enc (C, d) | None, [(id, start, end, conf), ...])"""
        pcm = np.ascontiguousarray(pcm, np.float32)
        cap_f, cap_e, cap_t = 4 + len(pcm) // 160, 2 + len(pcm) // 1280 + 2, 4096
        feats = np.zeros((cap_f, self.cfg.mel_bins), np.float32)
        enc = np.zeros((cap_e, self.cfg.d_model), np.float32)
        tok = np.zeros((cap_t, 3), np.int32)
        conf = np.zeros(cap_t, np.float32)
        nf, ne, nt = C.c_int(), C.c_int(), C.c_int()
        rc = self.L.pkref_stream_chunk(self.h, pcm.ctypes.data, len(pcm), feats.ctypes.data, cap_f, C.byref(nf), enc.ctypes.data,
                                       cap_e, C.byref(ne), tok.ctypes.data, conf.ctypes.data, cap_t, C.byref(nt))
        if rc != 0:
            raise RuntimeError("pkref_stream_chunk: " + self.L.pkref_stream_last_error().decode())
        toks = [(int(tok[i, 0]), int(tok[i, 1]), int(tok[i, 2]), float(conf[i])) for i in range(nt.value)]
        return (feats[:nf.value].copy() if nf.value else None, enc[:ne.value].copy() if ne.value else None, toks)


def _phrases(phrases):
    ids = np.array([t for ph in phrases for t in ph], np.int32)
    off = np.zeros(len(phrases) + 1, np.int32)
    off[1:] = np.cumsum([len(ph) for ph in phrases])
    return np.ascontiguousarray(ids if len(ids) else np.zeros(1, np.int32)), off


def ctc_greedy_boosted(lp, blank, phrases, boost=5.0):
    """ctc_greedy_decode_with_timestamps_boosted (src/phrase_boost.cpp:122-176) on one utterance."""
    L = lib()
    lp = np.ascontiguousarray(lp, np.float32)
    T, V = lp.shape
    ids, off = _phrases(phrases)
    o = [np.zeros(T, np.int32) for _ in range(3)]
    conf = np.zeros(T, np.float32)
    L.pkref_ctc_greedy_boosted.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    n = L.pkref_ctc_greedy_boosted(lp.ctypes.data, T, V, blank, ids.ctypes.data, off.ctypes.data, len(phrases), boost,
                                   o[0].ctypes.data, o[1].ctypes.data, o[2].ctypes.data, conf.ctypes.data)
    if n < 0:
        raise RuntimeError("pkref_ctc_greedy_boosted: " + L.pkref_last_error().decode())
    return [(int(o[0][i]), int(o[1][i]), int(o[2][i]), float(conf[i])) for i in range(n)]


def tdt_greedy_boosted(model, enc, phrases, boost=5.0, cap=8192):
    """tdt_greedy_decode_with_timestamps_boosted (src/phrase_boost.cpp:266-352) with a RefModel."""
    L = lib()
    enc = np.ascontiguousarray(enc, np.float32)
    ids, off = _phrases(phrases)
    o = [np.zeros(cap, np.int32) for _ in range(3)]
    conf = np.zeros(cap, np.float32)
    L.pkref_tdt_greedy_boosted.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    n = L.pkref_tdt_greedy_boosted(model.h, enc.ctypes.data, enc.shape[0], enc.shape[1], ids.ctypes.data, off.ctypes.data,
                                   len(phrases), boost, cap, o[0].ctypes.data, o[1].ctypes.data, o[2].ctypes.data, conf.ctypes.data)
    if n < 0:
        raise RuntimeError("pkref_tdt_greedy_boosted: " + L.pkref_last_error().decode())
    return [(int(o[0][i]), int(o[1][i]), int(o[2][i]), float(conf[i])) for i in range(min(n, cap))]


def tok_encode(model, text, cap=4096):
    L = lib()
    ids = np.zeros(cap, np.int32)
    L.pkref_tok_encode.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
    n = L.pkref_tok_encode(model.h, text.encode("utf-8"), cap, ids.ctypes.data)
    if n < 0:
        raise RuntimeError("pkref_tok_encode: " + L.pkref_last_error().decode())
    return ids[:n].tolist()


def resample(x, src_rate, dst_rate):
    """parakeet::resample (src/audio_io.cpp:238-251)."""
    L = lib()
    x = np.ascontiguousarray(x, np.float32)
    cap = int(len(x) * dst_rate / src_rate) + 16
    out = np.zeros(cap, np.float32)
    L.pkref_resample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    m = L.pkref_resample(x.ctypes.data, len(x), src_rate, dst_rate, out.ctypes.data, cap)
    if m < 0:
        raise RuntimeError("pkref_resample: " + L.pkref_last_error().decode())
    return out[:m].copy()
