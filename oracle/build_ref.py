"""oracle/_ref -- the UNMODIFIED reference (kratzert/RRMPG) installed beside the oracle.  TEST INFRASTRUCTURE.

    python oracle/build_ref.py            # in the build container, where /root/reference exists

`pip install --no-index --no-deps --target oracle/_ref` of the reference tree (from a scratch copy under /tmp,
because /root/reference is read-only and setuptools writes build/ and *.egg-info beside setup.py).  Nothing of it
is committed: `oracle/_ref/` is git-ignored, but not gpurun-ignored, so the installed package travels to the GPU
box like the built .so files.  The image on both sides carries numba 0.65, so `bench.py --impl reference` and the
`cpu_baseline` leg can time RRMPG's own numba path (`rrmpg.models.HBVEdu.simulate`, hbvedu.py:199-209 around
run_hbvedu, hbvedu_model.py:16) on the GPU box's host cores; where `oracle/_ref` or numba is missing they fall
back to the C port (oracle/rr_oracle.c) and say so (`kind: "port"`).

Only bench.py's reference / cpu_baseline legs and tests/ may import it (through `oracle.reference_package()`).
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("RRMPG_REFERENCE", "/root/reference")
TARGET = os.path.join(HERE, "_ref")


def build_ref(force=False):
    """Install the reference into oracle/_ref.  Returns the target path, or None when the reference tree is absent
    (the GPU box: the prebuilt directory travelled with the snapshot)."""
    if not os.path.isdir(os.path.join(REF_SRC, "rrmpg")):
        return TARGET if os.path.isdir(os.path.join(TARGET, "rrmpg")) else None
    if os.path.isdir(os.path.join(TARGET, "rrmpg")) and os.path.isdir(os.path.join(TARGET, "reference_tests")) and not force:
        return TARGET
    with tempfile.TemporaryDirectory(prefix="rrmpg_ref_") as tmp:
        src = os.path.join(tmp, "src")
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git", "docs", "examples"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation",
                               "--no-deps", "--find-links", "/opt/wheelhouse", "--target", TARGET, src])
        # the reference's own test-suite and its fixtures (test/test_models.py, test_tools.py, test_utils.py, test/data):
        # tests/test_reference_suite_gpu.py runs it VERBATIM against the drop-in package (rrmpg -> rrmpg_b200)
        shutil.copytree(os.path.join(REF_SRC, "test"), os.path.join(TARGET, "reference_tests"))
    return TARGET


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
