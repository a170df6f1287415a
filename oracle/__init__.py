"""ORACLE -- test infrastructure only.

CPU restatements of the reference's hot-path arithmetic, used as the checker by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  Nothing under
``gpv-1_b200/`` may import this package: the product path has no CPU fallback.

* ``matcher_oracle.c``  plain-C matcher cost + rectangular LSAP (pinned against the reference's
  ``HungarianMatcher`` and scipy by ``oracle/pin_matcher.py``).
* ``torch_oracle.py``   fp32 PyTorch restatement of the GPV-1 forward / criterion (pinned against the reference's
  own modules imported from /root/reference by ``oracle/make_golden.py``; fixtures in ``tests/golden``).
* ``ref_harness.py``    import shims that let the unmodified reference run in this container (only usable
  where /root/reference exists; never imported by tests that run on the GPU box).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (no FMA contraction: the reference rounds after every op)."""
    src = os.path.join(_HERE, "matcher_oracle.c")
    os.makedirs(_BUILD, exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_lsap_f64.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def matcher_cost(logits, boxes, tgt_boxes, tgt_labels, tgt_offsets, w_class=1.0, w_bbox=5.0, w_giou=2.0):
    """logits [B,Q,C] f32, boxes [B,Q,4] f32, ragged targets -> cost blocks [B,Q,Tmax] f32 (unused entries 0)."""
    logits = np.ascontiguousarray(logits, np.float32)
    boxes = np.ascontiguousarray(boxes, np.float32)
    tgt_boxes = np.ascontiguousarray(tgt_boxes, np.float32).reshape(-1, 4)
    tgt_labels = np.ascontiguousarray(tgt_labels, np.int64)
    tgt_offsets = np.ascontiguousarray(tgt_offsets, np.int32)
    B, Q, C = logits.shape
    sizes = np.diff(tgt_offsets)
    Tmax = int(sizes.max()) if B else 0
    cost = np.zeros((B, Q, max(Tmax, 0)), np.float32)
    if B and Tmax:
        lib().oracle_matcher_cost(_p(logits), _p(boxes), _p(tgt_boxes), _p(tgt_labels), _p(tgt_offsets), B, Q, C, Tmax,
                                  ctypes.c_float(w_class), ctypes.c_float(w_bbox), ctypes.c_float(w_giou), _p(cost))
    return cost


def lsap_batched(cost, tgt_offsets):
    """cost [B,Q,Tmax] f32 -> (idx_q, idx_t) int64 [B, min(Q,Tmax)], -1 padded, rows sorted by query."""
    cost = np.ascontiguousarray(cost, np.float32)
    tgt_offsets = np.ascontiguousarray(tgt_offsets, np.int32)
    B, Q, Tmax = cost.shape
    K = min(Q, Tmax)
    oq = np.full((B, K), -1, np.int64)
    ot = np.full((B, K), -1, np.int64)
    if B and K:
        lib().oracle_lsap(_p(cost), _p(tgt_offsets), B, Q, Tmax, _p(oq), _p(ot))
    return oq, ot


def lsap(cost2d):
    """Single rectangular problem in float64, scipy.optimize.linear_sum_assignment semantics."""
    c = np.ascontiguousarray(cost2d, np.float64)
    nr, nc = c.shape
    k = min(nr, nc)
    a = np.zeros(k, np.int64)
    b = np.zeros(k, np.int64)
    n = lib().oracle_lsap_f64(nr, nc, _p(c), _p(a), _p(b))
    if n < 0:
        raise ValueError("cost matrix is infeasible")
    return a[:n], b[:n]
