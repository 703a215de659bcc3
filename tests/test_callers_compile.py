"""the reference's own callers of the boundary -- its test program, the three MEX wrappers and the JNI glue --
compile UNCHANGED against this repository's include/ (stub mex.h / jni.h / tiffio.h / bzlib.h under tests/stubs: no Matlab, JDK or
libtiff in this image), and every library symbol they reference is exported by liblfm_b200.so.  Needs /root/reference (CPU box)."""
import os
import subprocess

import pytest

from conftest import ROOT

REF = "/root/reference"
CALLERS = ["test/mainTest_lfmIO.cxx", "matlabWrapper/writeLFMstack.cpp", "matlabWrapper/readLFMstack.cpp", "matlabWrapper/readLFMheader.cpp",
           "src/jni/org_janelia_simview_lfm_LFMJNI.cpp"]
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="the reference tree is not on this box")


def _exported():
    so = os.path.join(ROOT, "lightfieldmicroscopy_pc-bzip2_b200", "liblfm_b200.so")
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


@pytest.mark.parametrize("src", CALLERS)
def test_reference_caller_compiles_and_links_against_the_boundary(src, tmp_path):
    obj = str(tmp_path / "caller.o")
    cmd = ["g++", "-std=c++14", "-w", "-fPIC", "-c", os.path.join(REF, src), "-o", obj,
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(REF, "src", "jni")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, "%s does not compile against include/:\n%s" % (src, r.stderr[-3000:])
    undef = subprocess.run(["nm", "-u", obj], capture_output=True, text=True, check=True).stdout.split()
    undef = [u for u in undef if u not in ("U", "w")]
    ours = [u for u in undef if "klb_" in u or u in ("writeKLBstack", "writeKLBstackSlices", "readKLBheader", "readKLBstack",
                                                       "readKLBstackInPlace", "readKLBroiInPlace")]
    assert ours, "%s references nothing of the boundary?" % src
    missing = sorted(set(ours) - _exported())
    assert not missing, "%s needs symbols liblfm_b200.so does not export: %s" % (src, missing)
