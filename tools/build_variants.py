"""Build tuning variants of the library for A/B timing on the GPU box (developer tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor
from odevio_b200.build import build_library
VARIANTS = {
    "h3timeline": ["ODEVIO_H3_TIMELINE=1"],
    "h3cb0": ["H3_COMMIT_BATCH=0"],
    "h3skip0": ["H3_SKIP_LAST_BARRIER=0"],
    "h3noalias": ["H3_XJ_ALIAS=0"],
    "h3pair": ["H3_TMEM_PAIR=1"],
    "h3pair_cb0": ["H3_TMEM_PAIR=1", "H3_COMMIT_BATCH=0"],
    "h3tl5f1": ["ODEVIO_H3_TIMELINE=1", "H3_TL_STAGE=5", "H3_FUSE_STAGE_ARG=1"],
    "h3tl5f0": ["ODEVIO_H3_TIMELINE=1", "H3_TL_STAGE=5", "H3_FUSE_STAGE_ARG=0"],
    "w0p0": ["ODEVIO_PRODUCER_WAIT=0", "ODEVIO_EARLY_PROBE=0"],
    "w0p1": ["ODEVIO_PRODUCER_WAIT=0", "ODEVIO_EARLY_PROBE=1"],
    "w1p1": ["ODEVIO_PRODUCER_WAIT=1", "ODEVIO_EARLY_PROBE=1"],
    "w2p1": ["ODEVIO_PRODUCER_WAIT=2", "ODEVIO_EARLY_PROBE=1"],
}
names = sys.argv[1:] or list(VARIANTS)
with ThreadPoolExecutor(4) as ex:
    for path in ex.map(lambda n: build_library(variant=n, defines=VARIANTS[n]), names):
        print(path)
