"""Utterance-sharded data parallelism: the only multi-GPU strategy this
path needs.  Utterances are independent (per-utterance mel statistics, reference
src/audio.cpp:139-152; per-utterance decode state, src/tdt.cpp:46-58), so clip i goes to rank
i // ceil(N / world) (contiguous blocks), every rank runs mel -> encoder -> decode on its
block with replicated weights, and ONE collective -- an all-gather of fixed-stride int32
token rows (len, ids...) -- assembles the global result.  No data-path collective otherwise.

torch.distributed is plumbing here: backend "nccl" over NVLink on GPUs, "gloo" in the CPU
tests (tests/test_dist.py, world_size 2).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; blocks of ceil(n/world), last ranks may be short/empty."""
    per = -(-n_items // world)
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def rows_from_tokens(tokens: Sequence[Sequence[int]], cap: int) -> np.ndarray:
    """[(len, ids...)] int32 rows of width 1 + cap (the engine's device token buffer layout)."""
    out = np.zeros((len(tokens), 1 + cap), np.int32)
    for i, t in enumerate(tokens):
        if len(t) > cap:
            raise ValueError("token row exceeds capacity")
        out[i, 0] = len(t)
        out[i, 1:1 + len(t)] = t
    return out


def tokens_from_rows(rows: np.ndarray) -> List[List[int]]:
    return [[int(x) for x in r[1:1 + int(r[0])]] for r in rows]


def all_gather_rows(local_rows, n_items: int, world: int, rank: int, group=None):
    """The single exchange step.  local_rows: torch int32 tensor [n_local, W] (CPU for gloo, CUDA
    for nccl).  Every rank contributes ceil(n/world) rows (zero padded), one
    all_gather_into_tensor, then the padding is dropped.  Returns [n_items, W] on every rank."""
    import torch
    import torch.distributed as dist
    per = -(-n_items // world)
    W = local_rows.shape[1]
    send = torch.zeros((per, W), dtype=torch.int32, device=local_rows.device)
    send[: local_rows.shape[0]] = local_rows
    recv = torch.empty((world * per, W), dtype=torch.int32, device=local_rows.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    keep = []
    for r in range(world):
        lo, hi = shard_range(n_items, world, r)
        keep.append(recv[r * per: r * per + (hi - lo)])
    return torch.cat(keep, dim=0)


def transcribe_sharded(transcribe_local, pcms: Sequence[np.ndarray], cap: int, world: int, rank: int, device="cpu"):
    """transcribe_local(list_of_pcm) -> list of token-id lists for this rank's block (micro-batched
    by the caller's engine).  Returns the token lists of ALL clips on every rank."""
    import torch
    lo, hi = shard_range(len(pcms), world, rank)
    local = transcribe_local(list(pcms[lo:hi])) if hi > lo else []
    rows = torch.from_numpy(rows_from_tokens(local, cap)).to(device)
    if world == 1:
        return tokens_from_rows(rows.cpu().numpy())
    allr = all_gather_rows(rows, len(pcms), world, rank)
    return tokens_from_rows(allr.cpu().numpy())
