#!/usr/bin/env python3
"""Build the REAL reference into oracle/_ref/ (test infrastructure; never shipped, never on the product path).

Sources are compiled from where they lie under /root/reference; nothing is copied into the repository.
Patched working copies live only in a throw-away directory under /tmp:
  * klb_imageIO.cpp: `return 0;` added at the end of unPredictor / unPredictor_space / unPredictor_angle
    (reference lines 1821, 1895, 1970 fall off the end of a non-void function -> gcc emits a trap);
  * common.h:19: LFM_PREDICTOR_WAY made overridable so the three compile-time "ways" can be built.

Outputs (all git-ignored, all travel with gpurun):
  oracle/_ref/libbz2ref.so            vendored bzip2 1.0.6, plain gcc
  oracle/_ref/liblfmref_cpu_way{0,1,2}.so   reference orchestrator + its CUDA kernels executed on the CPU
                                      through oracle/cuda_shim (kernel launch = nested loop) -> runs without a GPU
  oracle/_ref/liblfmref_gpu_way{0,1,2}.so   the same sources through nvcc -arch=sm_100 (needs a GPU to run);
                                      used for the timed reference arm of bench.py.  (--gpu)
"""
import os, re, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LFM_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src")
BZ = os.path.join(SRC, "external", "bzip2-1.0.6")
OUT = os.path.join(HERE, "_ref")

def run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-4000:] + "\n")
        raise SystemExit(1)

def patch_imageio(text):
    # add the missing `return 0;` (three functions end with cudaFree(dpBuffer); cudaFree(dpSymbols); })
    pat = re.compile(r"(\tcudaFree\(dpBuffer\);\s*\n\tcudaFree\(dpSymbols\);\s*\n)(\})")
    text, n = pat.subn(r"\1\treturn 0;\n\2", text)
    assert n == 3, n
    return text

LAUNCH = re.compile(r"(\w+)\s*<<\s*<\s*(\w+)\s*,\s*(\w+)\s*>>\s*>\s*\(([^;]*)\)\s*;")

def build(gpu=False, force=False):
    if not os.path.isdir(SRC):
        print("build_ref: %s not present; keeping prebuilt oracle/_ref" % SRC)
        return False
    os.makedirs(OUT, exist_ok=True)
    want = ["libbz2ref.so"] + ["liblfmref_cpu_way%d.so" % w for w in range(3)]
    if gpu:
        want += ["liblfmref_gpu_way%d.so" % w for w in range(3)]
    if not force and all(os.path.exists(os.path.join(OUT, f)) for f in want):
        return True
    tmp = tempfile.mkdtemp(prefix="lfmref_build_")
    try:
        # ---- vendored bzip2 ----
        bzsrc = [os.path.join(BZ, f + ".c") for f in ("blocksort", "huffman", "crctable", "randtable", "compress", "decompress", "bzlib")]
        run(["gcc", "-O2", "-w", "-shared", "-fPIC", "-o", os.path.join(OUT, "libbz2ref.so")] + bzsrc)
        # ---- patched working copies (tmp only) ----
        w = os.path.join(tmp, "src"); os.makedirs(w)
        for f in os.listdir(SRC):
            p = os.path.join(SRC, f)
            if os.path.isfile(p) and f.split(".")[-1] in ("cpp", "h", "cu"):
                shutil.copy(p, os.path.join(w, f))
                os.chmod(os.path.join(w, f), 0o644)
        t = open(os.path.join(w, "klb_imageIO.cpp"), encoding="latin-1").read()
        open(os.path.join(w, "klb_imageIO.cpp"), "w", encoding="latin-1").write(patch_imageio(t))
        t = open(os.path.join(w, "common.h"), encoding="latin-1").read()
        t2 = t.replace("#define LFM_PREDICTOR_WAY (0)", "#ifndef LFM_PREDICTOR_WAY\n#define LFM_PREDICTOR_WAY (0)\n#endif")
        assert t2 != t
        open(os.path.join(w, "common.h"), "w", encoding="latin-1").write(t2)
        cus = ["lfm_Predictors", "lfm_Predictors_space", "lfm_Predictors_angle"]
        cpps = ["klb_imageHeader", "klb_ROI", "klb_circularDequeue", "klb_Cwrapper"]
        stubs = os.path.join(HERE, "stubs")
        # ---- CPU build through the shim ----
        shim = os.path.join(HERE, "cuda_shim")
        cw = os.path.join(tmp, "cpu"); os.makedirs(cw)
        inc = ["-I", shim, "-I", stubs, "-I", w, "-I", BZ]
        cxx = ["g++", "-O2", "-w", "-std=c++17", "-fPIC"]
        objs = []
        for c in cus:
            t = open(os.path.join(w, c + ".cu"), encoding="latin-1").read()
            t, n = LAUNCH.subn(r"LFMSHIM_LAUNCH(\1, \2, \3, \4);", t)
            assert n >= 7, (c, n)
            assert "<< <" not in t and "<<<" not in t
            open(os.path.join(cw, c + "_shim.cpp"), "w", encoding="latin-1").write(t)
            run(cxx + inc + ["-c", os.path.join(cw, c + "_shim.cpp"), "-o", os.path.join(cw, c + ".o")])
            objs.append(os.path.join(cw, c + ".o"))
        for c in cpps:
            run(cxx + inc + ["-c", os.path.join(w, c + ".cpp"), "-o", os.path.join(cw, c + ".o")])
            objs.append(os.path.join(cw, c + ".o"))
        for way in range(3):
            d = ["-DLFM_PREDICTOR_WAY=%d" % way]
            run(cxx + inc + d + ["-c", os.path.join(w, "klb_imageIO.cpp"), "-o", os.path.join(cw, "io%d.o" % way)])
            run(cxx + inc + d + ["-c", os.path.join(HERE, "ref_driver.cpp"), "-o", os.path.join(cw, "drv%d.o" % way)])
            run(["g++", "-shared", "-o", os.path.join(OUT, "liblfmref_cpu_way%d.so" % way),
                 os.path.join(cw, "io%d.o" % way), os.path.join(cw, "drv%d.o" % way)] + objs +
                ["-L", OUT, "-lbz2ref", "-lz", "-lpthread", "-Wl,-rpath,$ORIGIN"])
        # ---- real CUDA build (runs only on a GPU box) ----
        if gpu:
            gw = os.path.join(tmp, "gpu"); os.makedirs(gw)
            cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
            ginc = ["-I", stubs, "-I", w, "-I", BZ, "-I", os.path.join(cuda, "include")]
            gobjs = []
            for c in cus:
                run(["nvcc", "-O2", "-w", "-std=c++14", "-arch=sm_100", "-Xcompiler", "-fPIC"] + ginc +
                    ["-c", os.path.join(w, c + ".cu"), "-o", os.path.join(gw, c + ".o")])
                gobjs.append(os.path.join(gw, c + ".o"))
            gxx = ["g++", "-O2", "-w", "-std=c++14", "-fPIC"]
            for c in cpps:
                run(gxx + ginc + ["-c", os.path.join(w, c + ".cpp"), "-o", os.path.join(gw, c + ".o")])
                gobjs.append(os.path.join(gw, c + ".o"))
            for way in range(3):
                d = ["-DLFM_PREDICTOR_WAY=%d" % way]
                run(gxx + ginc + d + ["-c", os.path.join(w, "klb_imageIO.cpp"), "-o", os.path.join(gw, "io%d.o" % way)])
                run(gxx + ginc + d + ["-c", os.path.join(HERE, "ref_driver.cpp"), "-o", os.path.join(gw, "drv%d.o" % way)])
                run(["g++", "-shared", "-o", os.path.join(OUT, "liblfmref_gpu_way%d.so" % way),
                     os.path.join(gw, "io%d.o" % way), os.path.join(gw, "drv%d.o" % way)] + gobjs +
                    ["-L", OUT, "-lbz2ref", "-lz", "-lpthread", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                     "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + os.path.join(cuda, "lib64")])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return True

if __name__ == "__main__":
    ok = build(gpu="--gpu" in sys.argv, force="--force" in sys.argv)
    print("oracle/_ref:", sorted(os.listdir(OUT)) if os.path.isdir(OUT) else None)
