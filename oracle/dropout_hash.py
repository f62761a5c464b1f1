"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

numpy restatement of the counter-based dropout mask of fairmultimodal_b200/csrc/dropout.cuh, so that the masks the
CUDA kernels generate can be checked BIT-EXACTLY (integer work) instead of only statistically.  The reference draws its
masks from torch's Philox stream (nn.Dropout / SDPA dropout_p; HF modeling_bert.py:111,205,297,355; 10_FAME.py:214,255),
which no other implementation can reproduce; what this file pins is our own published mask function:

    seed'   = mix32(seed + step * 0x632BE5AB)
    rowseed = mix32(seed' ^ (row * 0x9E3779B9))
    unit    = column >> group_shift
    bits    = mix32(rowseed + (unit >> 1) * 0x85EBCA6B)
    keep    = (unit & 1 ? bits >> 16 : bits & 0xffff) >= thresh16
"""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def mix32(h):
    """lowbias32 integer finalizer on uint32 arrays (computed in uint64 to avoid numpy overflow warnings)."""
    h = np.asarray(h, dtype=np.uint64) & M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x7FEB352D)) & M32
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x846CA68B)) & M32
    h ^= h >> np.uint64(16)
    return h


def keep_mask(seed, step, rows, cols, thresh16, group_shift=0):
    """bool [rows, cols]: True where the element is kept."""
    site = mix32((np.uint64(seed) + np.uint64(step) * np.uint64(0x632BE5AB)) & M32)
    r = np.arange(rows, dtype=np.uint64)
    rowseed = mix32(site ^ ((r * np.uint64(0x9E3779B9)) & M32))                       # [rows]
    unit = np.arange(cols, dtype=np.uint64) >> np.uint64(group_shift)                 # [cols]
    bits = mix32((rowseed[:, None] + ((unit >> np.uint64(1)) * np.uint64(0x85EBCA6B) & M32)[None, :]) & M32)
    field = np.where((unit & np.uint64(1))[None, :] == 1, bits >> np.uint64(16), bits & np.uint64(0xFFFF))
    return field >= np.uint64(thresh16)


def inv_keep(thresh16):
    return np.float32(65536.0) / (np.float32(65536.0) - np.float32(thresh16))
