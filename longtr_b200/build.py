"""In-tree build of liblongtr_b200.so (CUDA kernels + C ABI + host mirror) for sm_100a.

    python -m longtr_b200.build [--force]

nvcc cross-compiles without a GPU.  The library lands next to the sources
(longtr_b200/csrc/liblongtr_b200.so) so that it travels with the repository snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OUT = os.path.join(CSRC, "liblongtr_b200.so")
SYNTH_DIR = os.path.join(HERE, "synth")
SYNTH_OUT = os.path.join(SYNTH_DIR, "libltr_synth.so")  # workload generator of configs 3 / 4: its own library, no product code
OBJ = os.path.join(CSRC, "build")

CU_SOURCES = ["viterbi_kernels.cu", "band_kernel.cu", "plan_kernels.cu", "posterior_kernel.cu", "stutter_kernel.cu", "abi.cu", "stutter_abi.cu",
              "edit_kernel.cu", "edit_abi.cu", "em_kernel.cu", "microbench.cu"]
CPP_SOURCES = ["host/flat_api.cpp", "host/host_types.cpp", "host/hap_aligner.cpp", "host/stutter_host.cpp",
               "host/genotyper.cpp", "host/pipeline.cpp", "host/locus_batcher.cpp", "host/bam_reader.cpp", "host/region_loader.cpp", "host/candidate_alleles.cpp", "host/poa.cpp", "host/fasta_reader.cpp", "host/vcf_writer.cpp", "host/region_pipeline.cpp", "synth_stutter.cpp"]
HEADERS = ["viterbi_core.cuh", "band_core.cuh", "viterbi_host.h", "kernels.h", "stutter_core.cuh", "ctx.h", "plan_device.cuh", "edit_core.cuh"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "--threads", "8",
              "-I" + INCLUDE, "-I" + CSRC]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers():
    hs = [os.path.join(CSRC, h) for h in HEADERS]
    hs += [os.path.join(INCLUDE, h) for h in os.listdir(INCLUDE)]
    hostdir = os.path.join(CSRC, "host")
    hs += [os.path.join(hostdir, h) for h in os.listdir(hostdir) if h.endswith(".h")]
    return hs


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs = []
    objs = []
    for src in CU_SOURCES + CPP_SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src.replace("/", "_") + ".o")
        objs.append(obj)
        if force or _stale(obj, [path] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", path, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    if jobs or force or not os.path.exists(OUT):
        run([NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-lz"])
    synth_src = [os.path.join(SYNTH_DIR, "synth.cpp"), os.path.join(SYNTH_DIR, "longtr_synth.h")]
    if force or _stale(SYNTH_OUT, synth_src + [os.path.join(INCLUDE, "longtr_b200.h")]):
        run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I" + INCLUDE,
             "-I" + SYNTH_DIR, "-o", SYNTH_OUT, synth_src[0]])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
