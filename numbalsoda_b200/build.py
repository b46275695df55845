"""Build libb200ode.so in-tree (numbalsoda_b200/libb200ode.so).

  1. nvcc compiles every solver kernel for sm_100a to LTO-IR fatbins, one per
     system size (-DB200_NEQ=n), and the built-in right-hand sides;
  2. the fatbins are embedded into the shared library (.incbin);
  3. g++ compiles the host side (CUDA driver API through dlopen, nvJitLink linked
     statically so that a different libnvJitLink already loaded in the process --
     torch bundles 12.8 -- cannot be picked up by mistake).

Run:  python -m numbalsoda_b200.build [--force]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libb200ode.so")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA, "bin", "nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=lto_100a"]
NVCC_FLAGS = ["-O3", "-lineinfo", "-rdc=true", "-fatbin", "-std=c++17"]

# __launch_bounds__(128, MINBLOCKS): register budget per system size (blocks of 128
# threads that must fit one SM); chosen so that ptxas does not spill (checked by
# tests/test_abi.py on the linked cubin).
MINBLOCKS = {"dop853": {1: 4, 2: 4, 3: 3, 4: 2, 5: 2, 6: 1, 7: 1, 8: 1},
             # (lsoda n <= 3: four blocks = 128 registers with 4 doubles spilled beat three blocks at
             # 161 registers by 4.7 % on Robertson -- 9.08 M against 8.68 M traj/s,
             # profiles/r2_sweep_lsoda_occupancy.log; 120 registers or 5 blocks lose)
             "lsoda": {1: 4, 2: 4, 3: 4, 4: 2, 5: 4, 6: 3, 7: 2, 8: 2},
             "ivp": {1: 4, 2: 4, 3: 3, 4: 2, 5: 2, 6: 1, 7: 1, 8: 1}}
# lsoda keeps (13 n + n^2) doubles per thread in shared memory (lsoda_kernel.cu):
# blocks of 64 threads from n = 5 so that several blocks still fit the 227 KB of an SM
THREADS = {"dop853": {n: 128 for n in range(1, 9)},
           "lsoda": {1: 128, 2: 128, 3: 128, 4: 128, 5: 64, 6: 64, 7: 64, 8: 64},
           "ivp": {n: 128 for n in range(1, 9)}}
# DOP853 dense-output queue: records per warp (dop853_kernel.cu); 64 = single inlined
# drain site; sized so that MINBLOCKS blocks fit the 227 KB of shared memory.
QCAP = {1: 55, 2: 55, 3: 55, 4: 63, 5: 63, 6: 63, 7: 63, 8: 63}


def _smem_ok(n, mb, qcap):
    return mb * 4 * (4 + 11 * n) * qcap * 8 <= 227 * 1024


for _n in QCAP:
    while not _smem_ok(_n, MINBLOCKS["dop853"][_n], QCAP[_n]) and QCAP[_n] > 36:
        QCAP[_n] -= 4

SOLVERS = {
    # name: (source, sizes)
    "dop853": ("dop853_kernel.cu", list(range(1, 9))),
    "lsoda": ("lsoda_kernel.cu", list(range(1, 9))),
    # solve_ivp: DOP853 with dense output after every step and terminal events
    "ivp": ("solve_ivp_kernel.cu", list(range(1, 9))),
}
BUILTIN_RHS = ["lotka_volterra", "lotka_volterra2", "lorenz", "robertson", "brusselator1d", "brusselator1d_il"]
BUILTIN_EVENTS = ["none", "lv_prey", "lorenz_z"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    images = []
    for name, (src, sizes) in SOLVERS.items():
        srcp = os.path.join(CSRC, src)
        if not os.path.exists(srcp):
            continue
        for n in sizes:
            out = os.path.join(OBJ, "%s_n%d.fatbin" % (name, n))
            # B200ODE_REBUILD="dop853_n3,lsoda_n3": rebuild exactly these images (launch-geometry
            # sweeps on the GPU box, tools/sweep_build.sh) and leave the others as they are
            only = os.environ.get("B200ODE_REBUILD")
            if only is not None:
                stale = ("%s_n%d" % (name, n)) in only.split(",")
            else:
                stale = force or _newer(out, [srcp, os.path.abspath(__file__)] + headers)
            if stale:
                mb = int(os.environ.get("B200ODE_MINBLOCKS", MINBLOCKS.get(name, {}).get(n, 1)))
                qc = int(os.environ.get("B200ODE_QCAP", QCAP.get(n, 48)))
                ni = int(os.environ.get("B200ODE_DRAIN_NOINLINE", "0"))
                th = int(os.environ.get("B200ODE_THREADS", THREADS.get(name, {}).get(n, 128)))
                # lsoda: pow out of line (one copy instead of five: +17 % on Robertson, the
                # kernel's largest stall reason is instruction-cache misses); out-of-line
                # division was measured slower (all sites 4.28 M, rare blocks only 4.40 M,
                # inline 4.66 M traj/s) and stays an experiment switch
                extra = []
                if name == "lsoda" and os.environ.get("B200ODE_POW_NOINLINE", "1") == "1":
                    extra.append("-DB200POW_NOINLINE")
                if name == "lsoda" and os.environ.get("B200ODE_DIV_NOINLINE", "0") == "1":
                    extra.append("-DB200_DIV_NOINLINE")
                if name == "lsoda":
                    extra += os.environ.get("B200ODE_LSODA_DEFS", "").split()
                if name == "dop853":
                    extra += os.environ.get("B200ODE_DOP853_DEFS", "").split()
                if name == "ivp":
                    extra += os.environ.get("B200ODE_IVP_DEFS", "").split()
                log = _run([NVCC] + ARCH + NVCC_FLAGS + extra + ["-DB200_NEQ=%d" % n, "-DB200_QCAP=%d" % qc,
                                                          "-DB200_MINBLOCKS=%d" % mb, "-DB200_DRAIN_NOINLINE=%d" % ni, "-DB200_THREADS=%d" % th,
                                                          "-o", out, srcp])
                if verbose and log.strip():
                    print(log)
            images.append(("%s_n%d" % (name, n), out))
    # LSODA, one warp per trajectory: a single image, system size at run time
    srcp = os.path.join(CSRC, "lsoda_warp_kernel.cu")
    out = os.path.join(OBJ, "lsodaw.fatbin")
    if force or _newer(out, [srcp, os.path.abspath(__file__)] + headers):
        _run([NVCC] + ARCH + NVCC_FLAGS + ["-DB200POW_NOINLINE"] +
             os.environ.get("B200ODE_EXTRA_DEFS", "").split() + ["-o", out, srcp])
    images.append(("lsodaw", out))
    # DOP853, one warp per trajectory
    srcp = os.path.join(CSRC, "dop853_warp_kernel.cu")
    out = os.path.join(OBJ, "dop853w.fatbin")
    if force or _newer(out, [srcp, os.path.abspath(__file__)] + headers):
        _run([NVCC] + ARCH + NVCC_FLAGS + ["-o", out, srcp])
    images.append(("dop853w", out))
    # solve_ivp, one warp per trajectory
    srcp = os.path.join(CSRC, "solve_ivp_warp_kernel.cu")
    out = os.path.join(OBJ, "ivpw.fatbin")
    if force or _newer(out, [srcp, os.path.abspath(__file__)] + headers):
        _run([NVCC] + ARCH + NVCC_FLAGS + ["-o", out, srcp])
    images.append(("ivpw", out))
    srcp = os.path.join(CSRC, "rhs_builtin.cu")
    for r in BUILTIN_RHS:
        out = os.path.join(OBJ, "rhs_%s.fatbin" % r)
        if force or _newer(out, [srcp]):
            _run([NVCC] + ARCH + NVCC_FLAGS + ["-DB200_RHS_%s" % r.upper(), "-o", out, srcp])
        images.append(("rhs_" + r, out))

    srcp = os.path.join(CSRC, "events_builtin.cu")
    for r in BUILTIN_EVENTS:
        out = os.path.join(OBJ, "events_%s.fatbin" % r)
        if force or _newer(out, [srcp]):
            _run([NVCC] + ARCH + NVCC_FLAGS + ["-DB200_EVENTS_%s" % r.upper(), "-o", out, srcp])
        images.append(("events_" + r, out))

    # default of the optional lane-parallel right-hand side (rhs_lanes_default.cu)
    srcp = os.path.join(CSRC, "rhs_lanes_default.cu")
    out = os.path.join(OBJ, "rhs_lanes_default.fatbin")
    if force or _newer(out, [srcp]):
        _run([NVCC] + ARCH + NVCC_FLAGS + ["-o", out, srcp])
    images.append(("rhs_lanes_default", out))

    # the handle's link flags as a link-time constant (link_flags.cu)
    srcp = os.path.join(CSRC, "link_flags.cu")
    for k in range(4):
        out = os.path.join(OBJ, "flags_%d.fatbin" % k)
        if force or _newer(out, [srcp]):
            _run([NVCC] + ARCH + NVCC_FLAGS + ["-DB200_LINK_FLAGS=%du" % k, "-o", out, srcp])
        images.append(("flags_%d" % k, out))

    # self-test kernels: a plain sm_100a cubin (no LTO link needed)
    srcp = os.path.join(CSRC, "selftest_kernel.cu")
    out = os.path.join(OBJ, "selftest.cubin")
    if force or _newer(out, [srcp] + headers):
        _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-cubin",
              "-std=c++17", "-o", out, srcp])
    images.append(("selftest", out))

    asm = os.path.join(OBJ, "embedded_images.S")
    lines = ['    .section .rodata', '']
    for i, (name, path) in enumerate(images):
        lines += ['    .balign 16', 'b200img_%d_begin:' % i, '    .incbin "%s"' % path,
                  'b200img_%d_end:' % i, 'b200img_%d_name:' % i, '    .asciz "%s"' % name, '']
    lines += ['    .section .data.rel.ro,"aw"', '    .balign 8',
              '    .globl b200ode_embedded_images', 'b200ode_embedded_images:']
    for i in range(len(images)):
        lines += ['    .quad b200img_%d_name' % i, '    .quad b200img_%d_begin' % i,
                  '    .quad b200img_%d_end' % i]
    lines += ['    .globl b200ode_embedded_count', '    .balign 4', 'b200ode_embedded_count:',
              '    .long %d' % len(images), '    .section .note.GNU-stack,"",@progbits', '']
    new_asm = "\n".join(lines)
    if not os.path.exists(asm) or open(asm).read() != new_asm:
        open(asm, "w").write(new_asm)

    cxx = shutil.which("g++") or "g++"
    deps = [p for _, p in images] + [asm, os.path.join(CSRC, "b200ode.cpp"),
                                     os.path.join(CSRC, "b200ode_launch.h"),
                                     os.path.join(os.path.dirname(HERE), "include", "b200ode.h")]
    if force or _newer(LIB, deps):
        cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-s", "-Wl,--gc-sections",
               "-Wall", "-Wextra"] + os.environ.get("B200ODE_EXTRA_DEFS", "").split() + [
               "-I" + os.path.join(CUDA, "include"),
               "-o", LIB, os.path.join(CSRC, "b200ode.cpp"), asm,
               "-Wl,--exclude-libs,ALL", "-L" + os.path.join(CUDA, "lib64"),
               "-l:libnvJitLink_static.a", "-l:libnvptxcompiler_static.a", "-ldl", "-lpthread", "-lrt"]
        log = _run(cmd)
        if verbose and log.strip():
            print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
