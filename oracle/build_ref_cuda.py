"""Build the UNMODIFIED reference CUDA op (`_msmv_sampling_cuda`) for sm_100a into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The sources are compiled where they lie under
/root/reference/models/csrc/msmv_sampling (never copied into this repo); only the
resulting `.so` lands in `oracle/_ref/` (git-ignored, but shipped to the GPU box by gpurun).
It serves two purposes on the GPU box:
  * oracle O3 of SURVEY.md §8(c): parity of our op against the reference kernel itself;
  * the "reference CUDA op" baseline the north-star's >=10x target is quoted against.
We do not run the reference's own build system (models/csrc/setup.py); this is our recipe:
torch.utils.cpp_extension.load() with an explicit -gencode for sm_100a.
"""
import os
import sys

REF_SRC = '/root/reference/models/csrc/msmv_sampling'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')


def build(verbose=False):
    if not os.path.isdir(REF_SRC):
        print('[build_ref_cuda] %s not present (GPU box?) - using prebuilt .so if any' % REF_SRC)
        return None
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, '_msmv_sampling_cuda.so')
    srcs = [os.path.join(REF_SRC, f) for f in
            ('msmv_sampling.cpp', 'msmv_sampling_forward.cu', 'msmv_sampling_backward.cu')]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    os.environ.setdefault('MAX_JOBS', '4')
    from torch.utils.cpp_extension import load
    load(name='_msmv_sampling_cuda', sources=srcs, extra_include_paths=[REF_SRC],
         extra_cuda_cflags=['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo'],
         build_directory=OUT, verbose=verbose, is_python_module=False)
    return so


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv))
