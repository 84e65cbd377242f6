"""Build an alternative libadmm_b200 with extra nvcc flags (kernel-tuning A/B on one box):
    python tools/build_variant.py out.so -DCOLS_BIG_C2160=4
Objects go to a scratch directory; the default library is untouched.  Use with ADMM_B200_LIB=$PWD/out.so."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import build as B
out, flags = sys.argv[1], sys.argv[2:]
objdir = "/tmp/admm_variant_%d" % os.getpid()
os.makedirs(objdir, exist_ok=True)
srcs = B._sources()
objs = [os.path.join(objdir, os.path.basename(s)[:-3] + ".o") for s in srcs]
def cc(so):
    r = subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *flags, "-c", so[0], "-o", so[1]], capture_output=True, text=True)
    if r.returncode: raise RuntimeError(r.stderr)
with ThreadPoolExecutor(8) as ex: list(ex.map(cc, zip(srcs, objs)))
subprocess.check_call([B._nvcc(), "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
print("built", out)
