"""Recipe for oracle/_ref: the UNMODIFIED reference package, made available to `bench.py --impl reference` on the GPU box
(where /root/reference does not exist).  Test / measurement infrastructure only; oracle/_ref is git-ignored (it still travels
with gpurun) and nothing under causaldiffae_b200/ ever imports it.

    python -m oracle.build_ref            # needs /root/reference (the build container)

Step 1 is the contract's own install line (`pip install --no-index --no-build-isolation --no-deps --target oracle/_ref`,
from a copy under /tmp because the reference tree is read-only).  The reference's setup.py declares
`py_modules=["improved_diffusion"]` for what is a package directory, so that wheel carries only metadata (observed: a 1.1 kB
wheel, no .py files) - upstream is meant to be used with `pip install -e .`.  Step 2 therefore places the package directory
itself (what an editable install would expose) next to the metadata.  The files are byte-identical to /root/reference: the
two stand-in modules (blobfile, mpi4py) and the four documented patches live in oracle/refshim.py, outside the copy."""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("CDAE_REFERENCE_ROOT", "/root/reference")


def build(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "improved_diffusion")):
        return None                                   # GPU box: use the prebuilt oracle/_ref as it travelled
    pkg = os.path.join(DST, "improved_diffusion")
    if os.path.isdir(pkg) and not filecmp.dircmp(os.path.join(SRC, "improved_diffusion"), pkg, ignore=["__pycache__"]).diff_files \
            and not filecmp.dircmp(os.path.join(SRC, "improved_diffusion"), pkg, ignore=["__pycache__"]).left_only:
        return DST
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns("__pycache__", "*.png"))
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
                            "--target", DST, work], capture_output=True, text=True)
        note = "pip install rc=%d" % r.returncode
        if verbose:
            print(note, r.stderr[-500:])
    if not os.path.isdir(pkg):                        # the metadata-only wheel described above
        shutil.copytree(os.path.join(SRC, "improved_diffusion"), pkg, ignore=shutil.ignore_patterns("__pycache__"))
        note += "; package directory placed by copy (setup.py lists it as a py_module)"
    with open(os.path.join(DST, "BUILD_NOTE.txt"), "w") as f:
        f.write(note + "\n")
    return DST


if __name__ == "__main__":
    print(build(verbose=True))
