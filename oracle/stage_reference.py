"""Stage the reference's pure-Python decoder modules for the GPU box -- TEST INFRASTRUCTURE.

The drop-in claim ("drops into the reference's forward pass unchanged") is proven by running the REAL reference modules
on the B200 with only `models.csrc.wrapper` swapped for `sparsebev_b200.wrapper` (tests/test_gpu_dropin.py).
/root/reference does not exist on the GPU box, so `build()` copies the five files that test imports

    models/sparsebev_transformer.py  models/sparsebev_sampling.py  models/utils.py  models/bbox/utils.py  models/checkpoint.py

byte for byte into `baseline/_ref/models/` -- git-ignored (never part of this repository's history), but not
gpurun-ignored, so it travels to the box exactly like oracle/_ref's compiled reference op does.  Nothing under
`sparsebev_b200/` reads that directory; when it is absent the drop-in tests skip.
"""
import os
import shutil

REF = '/root/reference/models'
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(os.path.dirname(HERE), 'baseline', '_ref', 'models')
FILES = ['sparsebev_transformer.py', 'sparsebev_sampling.py', 'utils.py', 'bbox/utils.py', 'checkpoint.py']


def stage():
    if not os.path.isdir(REF):
        return DST if os.path.isdir(DST) else None
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    return DST


if __name__ == '__main__':
    print(stage())
