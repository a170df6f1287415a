"""Developer tool (GPU box, under ncu): one launch of each variant of the row-tile-resident sub-layer kernels at the bench shape
(DETR encoder layer: B = 32, S = 300, d = 256, d_ff = 2048), reusing tools/prof_layer.py with its timing loop replaced by a single call.

    ncu --set full --clock-control none -k regex:'attn_block_fwd_kernel|mlp_block_fwd_kernel' -o /tmp/layer python tools/ncu_layer_once.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import prof_layer as pl  # noqa: E402


def once(fn, reps=1):
    fn()
    pl.torch.cuda.synchronize()
    return 1.0


pl.timeit = once
pl.attn(32, 300, 300, True)
pl.mlp(9600)
