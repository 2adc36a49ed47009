"""Token cache: the step BEFORE the hot path.

The reference's ``tools/dump_tokens.py:82-99`` writes the patch tokens of 35 images as one compressed
``.npz`` with keys ``tokens (n, N, C) float32``, ``images``, ``names`` -- a visualisation aid.  Training the
head at HBM speed needs the same ``(B, N, C)`` layout at data-set scale, so this module adds a shard format
that streams, and keeps the reference's file readable:

* ``write_shard`` / ``TokenShard``: one ``.eptok`` file = 4 KiB JSON header + raw bf16 tokens ``(n, N, C)``
  + int64 labels ``(n,)``, both 4 KiB-aligned so the payload can be memory-mapped and copied straight into
  pinned host memory (no parsing, no decompression on the training path).
* ``load_reference_npz``: reads a ``dump_tokens.py`` file (``tokens`` float32 -> bf16; ``images``/``names``
  passed through).
* ``TokenStream``: iterates fixed-size batches over a list of shards, rank-sharded like the reference's
  ``DistributedSampler`` + ``drop_last`` (main_linprobe.py:286-287,313-314), double-buffering the
  host->device copy on a side stream so batch i+1 is in flight while the trainer runs batch i; the batches
  are handed to ``EPHeadTrainer.train_step`` as device tensors.
"""
import json
import os
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

MAGIC = "EPTOK1"
HEADER_BYTES = 4096
ALIGN = 4096


def _align(n):
    return (n + ALIGN - 1) // ALIGN * ALIGN


def write_shard(path: str, tokens: torch.Tensor, labels: torch.Tensor, meta: Optional[dict] = None) -> None:
    """tokens (n, N, C) any float dtype (stored as bf16), labels (n,) integer."""
    if tokens.dim() != 3 or labels.shape != (tokens.shape[0],):
        raise ValueError("tokens must be (n, N, C) and labels (n,)")
    t = tokens.detach().to("cpu", torch.bfloat16).contiguous()
    y = labels.detach().to("cpu", torch.int64).contiguous()
    n, N, C = t.shape
    tok_off = HEADER_BYTES
    lab_off = tok_off + _align(t.numel() * 2)
    header = {"magic": MAGIC, "n": n, "N": N, "C": C, "dtype": "bfloat16", "tokens_offset": tok_off,
              "labels_offset": lab_off, "meta": meta or {}}
    raw = json.dumps(header).encode()
    if len(raw) > HEADER_BYTES:
        raise ValueError("shard metadata too large")
    with open(path, "wb") as f:
        f.write(raw.ljust(HEADER_BYTES, b"\0"))
        f.write(t.view(torch.int16).numpy().tobytes())
        f.seek(lab_off)
        f.write(y.numpy().tobytes())


class TokenShard:
    """A memory-mapped ``.eptok`` shard: ``tokens`` (n, N, C) bf16 and ``labels`` (n,) int64 views."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            header = json.loads(f.read(HEADER_BYTES).rstrip(b"\0").decode())
        if header.get("magic") != MAGIC:
            raise ValueError(f"{path}: not an {MAGIC} token shard")
        self.path, self.header = path, header
        self.n, self.N, self.C = header["n"], header["N"], header["C"]
        mm = np.memmap(path, dtype=np.uint8, mode="c")            # copy-on-write: torch wants a writable array, the file is never written
        tok = mm[header["tokens_offset"]: header["tokens_offset"] + self.n * self.N * self.C * 2]
        lab = mm[header["labels_offset"]: header["labels_offset"] + self.n * 8]
        self.tokens = torch.from_numpy(tok.view(np.int16).reshape(self.n, self.N, self.C)).view(torch.bfloat16)
        self.labels = torch.from_numpy(lab.view(np.int64).reshape(self.n))

    def __len__(self):
        return self.n


def load_reference_npz(path: str):
    """Read a ``tools/dump_tokens.py`` file: returns (tokens bf16 (n, N, C), images, names)."""
    z = np.load(path, allow_pickle=False)
    if "tokens" not in z.files:
        raise ValueError(f"{path}: no 'tokens' array (keys: {z.files})")
    tokens = torch.from_numpy(np.ascontiguousarray(z["tokens"])).to(torch.bfloat16)
    return tokens, (z["images"] if "images" in z.files else None), (z["names"] if "names" in z.files else None)


def epoch_order(total: int, epoch: int, seed: int, rank: int, world: int, batch: int, shuffle: bool = True):
    """Sample indices of this rank for one epoch: a seeded permutation (seed + epoch, as
    DistributedSampler.set_epoch does), strided over the ranks, truncated to whole batches (drop_last)."""
    g = torch.Generator().manual_seed(seed + epoch)
    order = torch.randperm(total, generator=g) if shuffle else torch.arange(total)
    order = order[: total - total % world][rank::world]
    return order[: len(order) - len(order) % batch]


class TokenStream:
    """Batches of cached tokens for one rank, staged through pinned memory onto the GPU ahead of use."""

    def __init__(self, shards: Sequence[TokenShard], batch: int, device, rank: int = 0, world: int = 1, seed: int = 0,
                 shuffle: bool = True, slots: int = 2):
        if not shards:
            raise ValueError("no shards")
        N, C = shards[0].N, shards[0].C
        if any(s.N != N or s.C != C for s in shards):
            raise ValueError("all shards must share (N, C)")
        self.shards, self.batch, self.device = list(shards), batch, torch.device(device)
        self.rank, self.world, self.seed, self.shuffle = rank, world, seed, shuffle
        self.N, self.C = N, C
        self.offsets = np.cumsum([0] + [len(s) for s in shards])
        self.total = int(self.offsets[-1])
        self.slots = slots
        self._hx = [torch.empty(batch, N, C, dtype=torch.bfloat16).pin_memory() for _ in range(slots)]
        self._hy = [torch.empty(batch, dtype=torch.int64).pin_memory() for _ in range(slots)]
        self._dx = [torch.empty(batch, N, C, dtype=torch.bfloat16, device=self.device) for _ in range(slots)]
        self._dy = [torch.empty(batch, dtype=torch.int64, device=self.device) for _ in range(slots)]
        self._copy = torch.cuda.Stream(device=self.device)
        self._ready = [torch.cuda.Event() for _ in range(slots)]
        self._free = [torch.cuda.Event() for _ in range(slots)]

    def steps_per_epoch(self) -> int:
        return (self.total - self.total % self.world) // self.world // self.batch

    def _gather(self, idx: torch.Tensor, slot: int):
        """Rows `idx` (global sample numbers) of the shards -> the slot's pinned buffers, one vectorised gather per
        shard touched (a per-sample Python loop costs ~20 us a row: 20 ms per 1024-sample batch, twice the H2D copy)."""
        hx, hy = self._hx[slot], self._hy[slot]
        idx_np = idx.numpy()
        shard_id = np.searchsorted(self.offsets, idx_np, side="right") - 1
        for s in np.unique(shard_id):
            sh = self.shards[int(s)]
            pos = np.nonzero(shard_id == s)[0]
            local = torch.from_numpy(idx_np[pos] - self.offsets[s])
            if len(pos) == len(idx_np):                    # the whole batch comes from this shard: gather in place
                torch.index_select(sh.tokens, 0, local, out=hx)
                torch.index_select(sh.labels, 0, local, out=hy)
            else:
                pos_t = torch.from_numpy(pos)
                hx.index_copy_(0, pos_t, sh.tokens.index_select(0, local))
                hy.index_copy_(0, pos_t, sh.labels.index_select(0, local))

    def _stage(self, idx: torch.Tensor, slot: int):
        self._free[slot].synchronize()                       # the consumer has finished with this slot
        self._gather(idx, slot)
        with torch.cuda.stream(self._copy):
            self._dx[slot].copy_(self._hx[slot], non_blocking=True)
            self._dy[slot].copy_(self._hy[slot], non_blocking=True)
            self._ready[slot].record(self._copy)

    def epoch(self, epoch: int) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Yields (tokens, labels) device tensors; each stays valid until the next-but-one batch is requested."""
        order = epoch_order(self.total, epoch, self.seed, self.rank, self.world, self.batch, self.shuffle)
        nb = len(order) // self.batch
        cur = torch.cuda.current_stream(self.device)
        for s in range(self.slots):
            self._free[s].record(cur)
        for b in range(min(self.slots - 1, nb)):
            self._stage(order[b * self.batch:(b + 1) * self.batch], b % self.slots)
        for b in range(nb):
            nxt = b + self.slots - 1
            if nxt < nb:
                self._stage(order[nxt * self.batch:(nxt + 1) * self.batch], nxt % self.slots)
            slot = b % self.slots
            cur.wait_event(self._ready[slot])
            yield self._dx[slot], self._dy[slot]
            self._free[slot].record(cur)
