"""Install the UNMODIFIED reference (georgegrosu1/torch-admm-deconv, /root/reference) into baseline/_ref.

    python baseline/install_ref.py

Recipe of the task contract: `pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
--target baseline/_ref <copy of /root/reference>` (a copy under /tmp, because /root/reference is read-only and setuptools
writes build files into the source tree; --no-deps because the reference pins torch==2.4.1+cu121 and friends, which
the image replaces with torch 2.11).  `baseline/_ref/` is git-ignored (no reference source in the history) but NOT
gpurun-ignored, so the installed package travels to the GPU box, where `bench.py --impl reference`, bench.py's
`cpu_baseline` leg and the end-to-end model test import `admmtor` from it.  Nothing in the product package does.
Outcome in this image: builds a pure-Python wheel admmtor-0.1.0 and installs it (recorded in DESIGN.md section 8).
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"


def installed():
    return os.path.exists(os.path.join(TARGET, "admmtor", "eops", "deconv.py"))


def install(force=False):
    """Returns the target directory, or None when the reference is not available (GPU box: uses the prebuilt copy)."""
    if installed() and not force:
        return TARGET
    if not os.path.isdir(REFERENCE):
        return None
    tmp = tempfile.mkdtemp(prefix="admmtor_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REFERENCE, src)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not installed():
            raise RuntimeError("pip install of the reference failed:\n" + r.stdout + r.stderr)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return TARGET


def import_reference():
    """Put baseline/_ref on sys.path and return (fft_admm_tv, ADMMDeconv) of the reference."""
    if not installed():
        raise ImportError("baseline/_ref is empty: run `python baseline/install_ref.py` in the build container")
    if TARGET not in sys.path:
        sys.path.insert(0, TARGET)
    from admmtor.eops.deconv import fft_admm_tv
    from admmtor.elayers.admmdeconv import ADMMDeconv
    return fft_admm_tv, ADMMDeconv


if __name__ == "__main__":
    print("installed at", install(force="--force" in sys.argv))
