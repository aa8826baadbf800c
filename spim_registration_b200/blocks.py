"""Block decomposition -- host-side mirror of the reference's ``Block`` / ``BlockGenerator`` classes.

Reference (under /root/reference/src/main/java/):
  spim/process/cuda/BlockGeneratorFixedSizePrecise.java:46-122   divideIntoBlocks (gen-2)
  spim/process/cuda/Block.java:134-251                           copyBlock / pasteBlock
  mpicbg/spim/postprocessing/deconvolution2/Block.java:376-458   divideIntoBlocks (gen-1, doubles too-small blocks)

Dims are given in the reference's (x, y, z) order at this API (like the Java classes); numpy
volumes are ``[z, y, x]``.  This is host logic (index arithmetic and memory copies), exactly as in
the reference, where blocks are cut on the CPU before each JNA call.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

EXT_ZERO, EXT_CONSTANT, EXT_MIRROR_SINGLE, EXT_MIRROR_DOUBLE, EXT_PERIODIC = 0, 1, 2, 3, 4


def _ext_index(a: np.ndarray, n: int, mode: int) -> np.ndarray:
    a = np.asarray(a, dtype=np.int64)
    if mode in (EXT_ZERO, EXT_CONSTANT):
        return np.where((a >= 0) & (a < n), a, -1)
    if mode == EXT_PERIODIC:
        return np.mod(a, n)
    if mode == EXT_MIRROR_SINGLE:
        if n == 1:
            return np.zeros_like(a)
        p = 2 * (n - 1)
        m = np.mod(a, p)
        return np.where(m < n, m, p - m)
    if mode == EXT_MIRROR_DOUBLE:
        p = 2 * n
        m = np.mod(a, p)
        return np.where(m < n, m, p - 1 - m)
    raise ValueError(f"unknown extension mode {mode}")


class Block:
    """``spim.process.cuda.Block`` (Block.java:34-131).  All tuples are (x, y, z)."""

    def __init__(self, blockSize, offset, effectiveSize, effectiveOffset, effectiveLocalOffset, isPrecise=True):
        self.blockSize = tuple(int(v) for v in blockSize)
        self.offset = tuple(int(v) for v in offset)
        self.effectiveSize = tuple(int(v) for v in effectiveSize)
        self.effectiveOffset = tuple(int(v) for v in effectiveOffset)
        self.effectiveLocalOffset = tuple(int(v) for v in effectiveLocalOffset)
        self.isPrecise = isPrecise

    def getBlockSize(self):
        return self.blockSize

    def getOffset(self):
        return self.offset

    def getEffectiveSize(self):
        return self.effectiveSize

    def getEffectiveOffset(self):
        return self.effectiveOffset

    def getEffectiveLocalOffset(self):
        return self.effectiveLocalOffset

    def copyBlock(self, source: np.ndarray, block: np.ndarray, ext: int = EXT_MIRROR_SINGLE, value: float = 0.0) -> None:
        """Block.java:134-215 -- fill ``block`` from ``source`` read through an out-of-bounds
        extension (``Views.extendMirrorSingle`` for conv1, ``Views.extendValue(1)`` for conv2)."""
        out = source
        for ax in range(3):                       # numpy axis ax <-> reference dim 2-ax
            d = 2 - ax
            coords = np.arange(self.offset[d], self.offset[d] + self.blockSize[d])
            idx = _ext_index(coords, source.shape[ax], ext)
            taken = np.take(out, np.where(idx < 0, 0, idx), axis=ax)
            if ext in (EXT_ZERO, EXT_CONSTANT):
                c = 0.0 if ext == EXT_ZERO else value
                shape = [1, 1, 1]
                shape[ax] = len(idx)
                taken = np.where((idx < 0).reshape(shape), np.float32(c), taken)
            out = taken
        block[...] = out

    def pasteBlock(self, target: np.ndarray, block: np.ndarray) -> None:
        """Block.java:217-251 -- write back the effective region only."""
        src = tuple(slice(self.effectiveLocalOffset[2 - ax], self.effectiveLocalOffset[2 - ax] + self.effectiveSize[2 - ax])
                    for ax in range(3))
        dst = tuple(slice(self.effectiveOffset[2 - ax], self.effectiveOffset[2 - ax] + self.effectiveSize[2 - ax])
                    for ax in range(3))
        target[dst] = block[src]


class BlockGeneratorFixedSizePrecise:
    """``spim.process.cuda.BlockGeneratorFixedSizePrecise`` (gen-2)."""

    def __init__(self, blockSize: Sequence[int]):
        self.blockSize = tuple(int(v) for v in blockSize)

    def divideIntoBlocks(self, imgSize: Sequence[int], kernelSize: Sequence[int]) -> Optional[List[Block]]:
        return divide_into_blocks(imgSize, self.blockSize, kernelSize, double_too_small=False)


def divide_into_blocks(imgSize: Sequence[int], blockSize: Sequence[int], kernelSize: Sequence[int],
                       double_too_small: bool = False) -> Optional[List[Block]]:
    """BlockGeneratorFixedSizePrecise.java:46-122; with ``double_too_small`` the gen-1 behaviour of
    D2/Block.java:389-395 (double the block size in a too-small dimension and retry)."""
    nd = len(imgSize)
    blockSize = [int(b) for b in blockSize]
    while True:
        eff = [blockSize[d] - int(kernelSize[d]) + 1 for d in range(nd)]
        if all(e > 0 for e in eff):
            break
        if not double_too_small:
            return None
        for d in range(nd):
            if eff[d] <= 0:
                blockSize[d] *= 2
    local = [int(kernelSize[d]) // 2 for d in range(nd)]
    num = [int(imgSize[d]) // eff[d] + (1 if int(imgSize[d]) % eff[d] else 0) for d in range(nd)]
    blocks: List[Block] = []
    # LocalizingZeroMinIntervalIterator: dimension 0 (x) fastest
    idx = [0] * nd
    total = int(np.prod(num))
    for _ in range(total):
        eff_off = [idx[d] * eff[d] for d in range(nd)]
        off = [eff_off[d] - int(kernelSize[d]) // 2 for d in range(nd)]
        eff_sz = [min(eff[d], int(imgSize[d]) - eff_off[d]) for d in range(nd)]
        blocks.append(Block(blockSize, off, eff_sz, eff_off, local, True))
        for d in range(nd):
            idx[d] += 1
            if idx[d] < num[d]:
                break
            idx[d] = 0
    return blocks
