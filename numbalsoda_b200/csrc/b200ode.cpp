// libb200ode.so -- host side of the C ABI declared in include/b200ode.h.
//
// What the reference does in-process on the CPU (one LSODA / dop853 object per
// call, src/wrapper.cpp:15, src/dop853_c_interface.f90:96) happens here on the GPU:
//   b200ode_link_rhs()        nvJitLink: solver LTO-IR (embedded, built by nvcc for
//                             sm_100a) + the user's RHS LTO-IR -> one cubin
//   b200ode_*_batched_device  cuLaunchKernel on caller-owned device buffers
//   b200ode_*_batched         the same behind host buffers (chunked, copies overlapped)
// The CUDA driver is loaded with dlopen at first use so that the library itself
// loads (and links RHS images) on machines without a GPU; there is no CPU solver
// in here: without a device every solve call fails with B200ODE_ENODEV.
#include "../../include/b200ode.h"

#include <cuda.h>
#include <dlfcn.h>
#include <nvJitLink.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "b200ode_launch.h"

// ---------------------------------------------------------------- embedded images
struct B200EmbeddedImage {
    const char* name;
    const unsigned char* begin;
    const unsigned char* end;
};
extern "C" const B200EmbeddedImage b200ode_embedded_images[];  // embedded_images.S/.c
extern "C" const int b200ode_embedded_count;

static const B200EmbeddedImage* find_image(const std::string& name)
{
    for (int i = 0; i < b200ode_embedded_count; i++)
        if (name == b200ode_embedded_images[i].name) return &b200ode_embedded_images[i];
    return nullptr;
}

// ---------------------------------------------------------------- errors
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...)
{
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
extern "C" const char* b200ode_last_error(void) { return g_err.c_str(); }
extern "C" int b200ode_version(void) { return B200ODE_VERSION; }

// ---------------------------------------------------------------- CUDA driver (dlopen)
#define DRV_FUNCS(X)                                                                        \
    X(cuInit) X(cuDeviceGet) X(cuDeviceGetCount) X(cuDeviceGetAttribute)                    \
    X(cuDevicePrimaryCtxRetain) X(cuDevicePrimaryCtxRelease_v2) X(cuCtxSetCurrent) X(cuCtxGetCurrent) X(cuCtxGetDevice)     \
    X(cuModuleLoadData) X(cuModuleUnload) X(cuModuleGetFunction) X(cuFuncGetAttribute) X(cuFuncSetAttribute)      \
    X(cuOccupancyMaxActiveBlocksPerMultiprocessor) X(cuLaunchKernel) X(cuMemAlloc_v2)       \
    X(cuMemFree_v2) X(cuMemcpyDtoH_v2) X(cuMemcpyHtoDAsync_v2) X(cuMemcpyDtoHAsync_v2) X(cuMemsetD8Async)      \
    X(cuStreamCreate) X(cuStreamDestroy_v2) X(cuStreamSynchronize) X(cuStreamWaitEvent)     \
    X(cuEventCreate) X(cuEventDestroy_v2) X(cuEventRecord) X(cuEventSynchronize)            \
    X(cuGetErrorString) X(cuMemGetInfo_v2) X(cuMemHostAlloc) X(cuMemFreeHost) X(cuMemcpyHtoD_v2)

struct Driver {
    void* lib = nullptr;
    bool ok = false;
#define X(f) decltype(&::f) f = nullptr;
    DRV_FUNCS(X)
#undef X
};
static Driver g_drv;
static std::once_flag g_drv_once;

static void load_driver()
{
    g_drv.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!g_drv.lib) return;
    bool all = true;
#define X(f)                                                      \
    g_drv.f = reinterpret_cast<decltype(&::f)>(dlsym(g_drv.lib, #f)); \
    if (!g_drv.f) all = false;
    DRV_FUNCS(X)
#undef X
    if (!all) return;
    if (g_drv.cuInit(0) != CUDA_SUCCESS) return;
    g_drv.ok = true;
}
static bool driver()
{
    std::call_once(g_drv_once, load_driver);
    return g_drv.ok;
}
static int cu_fail(CUresult r, const char* what)
{
    const char* s = nullptr;
    if (g_drv.cuGetErrorString) g_drv.cuGetErrorString(r, &s);
    return fail(B200ODE_ECUDA, "%s: CUDA error %d (%s)", what, (int)r, s ? s : "?");
}
#define CU(call)                                        \
    do {                                                \
        CUresult r_ = g_drv.call;                       \
        if (r_ != CUDA_SUCCESS) return cu_fail(r_, #call); \
    } while (0)

extern "C" int b200ode_device_count(void)
{
    if (!driver()) return 0;
    int n = 0;
    if (g_drv.cuDeviceGetCount(&n) != CUDA_SUCCESS) return 0;
    return n;
}

// ---------------------------------------------------------------- handle
static const int kCounterSlots = 64;

struct PerDevice {
    CUcontext ctx = nullptr;
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;      // parameters staged in registers (np >= 0)
    CUfunction fn_ptr = nullptr;  // data record passed by address (np < 0)
    CUfunction fn_vtol = nullptr, fn_ptr_vtol = nullptr;  // dop853 thread kernel: per-component tolerances
    CUfunction fn_band = nullptr; // warp lsoda: the same kernel built for 12 resident warps (banded matrix)
    CUfunction fn_tmem = nullptr; // warp lsoda, 32 < n <= 64: dense matrix in tensor memory, 8 warps per block
    CUdeviceptr next = 0;  // kCounterSlots work-queue heads
    int regs = 0, local_bytes = 0, smem = 0, blocks_per_sm = 0, sms = 0;
    int dyn_smem = 0;  // dynamic shared memory per block
    int max_dyn_smem = 0;  // what the kernels were allowed (CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES)
    int qcap = 0;      // dop853: records per warp queue
    // warp lsoda: per-block iteration matrices in global memory (L2 resident), one set per
    // launch that may be in flight; a set is reused only after its previous launch finished
    static const int kScratchSets = 4;
    CUdeviceptr wm_scratch = 0;
    size_t wm_set_bytes = 0;
    CUevent wm_done[kScratchSets] = {nullptr, nullptr, nullptr, nullptr};
    bool wm_used[kScratchSets] = {false, false, false, false};
    int threads = B200ODE_THREADS;  // threads per block the image was built for
};

struct b200ode_rhs_s {
    int solver = 0, neq = 0;
    int n_events = 0;  // solve_ivp handles
    unsigned flags = 0;
    std::vector<char> cubin;
    std::mutex mu;
    std::map<int, PerDevice> dev;
    std::atomic<unsigned> slot{0};
    std::atomic<int> launches{0};
    int last_grid = 0;
};

// Systems larger than B200ODE_THREAD_NMAX run one trajectory per warp
// (dop853_warp_kernel.cu, lsoda_warp_kernel.cu: one image each, size passed at run time).
static bool warp_kernel(int solver, int neq)
{
    (void)solver;
    return neq > B200ODE_THREAD_NMAX;
}
static const char* kernel_name(int solver, int neq)
{
    if (solver == B200ODE_SOLVE_IVP)
        return warp_kernel(solver, neq) ? "b200ode_ivpw_kernel" : "b200ode_ivp_kernel";
    if (solver == B200ODE_DOP853)
        return warp_kernel(solver, neq) ? "b200ode_dop853w_kernel" : "b200ode_dop853_kernel";
    return warp_kernel(solver, neq) ? "b200ode_lsodaw_kernel" : "b200ode_lsoda_kernel";
}

// One user image of a link: kind (B200ODE_IMG_*) -> what nvJitLink takes.
struct UserImage {
    const void* data = nullptr;
    size_t size = 0;
    nvJitLinkInputType type = NVJITLINK_INPUT_LTOIR;
};
static int resolve_image(const void* image, size_t size, int kind, const char* builtin_prefix,
                         const char* what, UserImage* out)
{
    if (!image) return fail(B200ODE_EINVAL, "%s image is NULL", what);
    out->data = image;
    out->size = size;
    switch (kind) {
    case B200ODE_IMG_LTOIR: out->type = NVJITLINK_INPUT_LTOIR; break;
    case B200ODE_IMG_PTX: out->type = NVJITLINK_INPUT_PTX; break;
    case B200ODE_IMG_FATBIN: out->type = NVJITLINK_INPUT_FATBIN; break;
    case B200ODE_IMG_BUILTIN: {
        std::string n = std::string(builtin_prefix) + static_cast<const char*>(image);
        const B200EmbeddedImage* r = find_image(n);
        if (!r) return fail(B200ODE_EINVAL, "unknown built-in %s '%s'", what,
                            static_cast<const char*>(image));
        out->data = r->begin;
        out->size = (size_t)(r->end - r->begin);
        out->type = NVJITLINK_INPUT_FATBIN;
        break;
    }
    default: return fail(B200ODE_EINVAL, "unknown image kind %d", kind);
    }
    if (out->size == 0) return fail(B200ODE_EINVAL, "empty %s image", what);
    return 0;
}

// nvJitLink: solver image + right-hand side (+ events function) + link flags -> cubin.
static int link_handle(int solver, int neq, const char* name, const UserImage& rhs,
                       const UserImage* events, int n_events, unsigned flags, b200ode_rhs_t* out,
                       const UserImage* lanes = nullptr, bool rhs_is_builtin = false)
{
    const B200EmbeddedImage* solver_img = find_image(name);
    if (!solver_img)
        return fail(B200ODE_ENOSYS, "no %s kernel was built for neq = %d (image %s missing)",
                    solver == B200ODE_DOP853 ? "dop853" : solver == B200ODE_LSODA ? "lsoda" : "solve_ivp",
                    neq, name);
    std::vector<const char*> opts = {"-arch=sm_100a", "-lto", "-O3", "-lineinfo"};
    if (flags & B200ODE_STRICT_FP) opts.push_back("-fma=0");
    nvJitLinkHandle h;
    nvJitLinkResult r = nvJitLinkCreate(&h, (uint32_t)opts.size(), opts.data());
    if (r != NVJITLINK_SUCCESS) return fail(B200ODE_ELINK, "nvJitLinkCreate failed (%d)", (int)r);
    auto log_and_fail = [&](const char* what, nvJitLinkResult rr) {
        size_t n = 0;
        std::string log;
        if (nvJitLinkGetErrorLogSize(h, &n) == NVJITLINK_SUCCESS && n > 1) {
            log.resize(n);
            nvJitLinkGetErrorLog(h, &log[0]);
        }
        nvJitLinkDestroy(&h);
        return fail(B200ODE_ELINK, "%s failed (%d): %s", what, (int)rr, log.c_str());
    };
    r = nvJitLinkAddData(h, NVJITLINK_INPUT_FATBIN, solver_img->begin,
                         (size_t)(solver_img->end - solver_img->begin), name);
    if (r != NVJITLINK_SUCCESS) return log_and_fail("nvJitLinkAddData(solver)", r);
    r = nvJitLinkAddData(h, rhs.type, rhs.data, rhs.size, "b200ode_user_rhs");
    if (r != NVJITLINK_SUCCESS) return log_and_fail("nvJitLinkAddData(rhs)", r);
    if (events) {
        r = nvJitLinkAddData(h, events->type, events->data, events->size, "b200ode_user_events");
        if (r != NVJITLINK_SUCCESS) return log_and_fail("nvJitLinkAddData(events)", r);
    }
    if (warp_kernel(solver, neq)) {
        // the lane-parallel form of the right-hand side (b200ode_warp.cuh): the caller's, the
        // built-in problem's own (its image defines both symbols), or the default that lets
        // lane 0 call the scalar function
        if (lanes) {
            r = nvJitLinkAddData(h, lanes->type, lanes->data, lanes->size, "b200ode_user_rhs_lanes");
            if (r != NVJITLINK_SUCCESS) return log_and_fail("nvJitLinkAddData(rhs_lanes)", r);
        } else if (!rhs_is_builtin) {
            const B200EmbeddedImage* d = find_image("rhs_lanes_default");
            if (!d) {
                nvJitLinkDestroy(&h);
                return fail(B200ODE_ENOSYS, "image rhs_lanes_default missing");
            }
            r = nvJitLinkAddData(h, NVJITLINK_INPUT_FATBIN, d->begin, (size_t)(d->end - d->begin), "rhs_lanes_default");
            if (r != NVJITLINK_SUCCESS) return log_and_fail("nvJitLinkAddData(rhs_lanes_default)", r);
        }
    }
    {
        // the flags as a device-side link-time constant (link_flags.cu)
        char fname[32];
        snprintf(fname, sizeof fname, "flags_%u", flags & 3u);
        const B200EmbeddedImage* f = find_image(fname);
        if (!f) {
            nvJitLinkDestroy(&h);
            return fail(B200ODE_ENOSYS, "image %s missing", fname);
        }
        r = nvJitLinkAddData(h, NVJITLINK_INPUT_FATBIN, f->begin, (size_t)(f->end - f->begin), fname);
        if (r != NVJITLINK_SUCCESS) return log_and_fail("nvJitLinkAddData(flags)", r);
    }
    r = nvJitLinkComplete(h);
    if (r != NVJITLINK_SUCCESS) return log_and_fail("nvJitLinkComplete", r);
    size_t n = 0;
    r = nvJitLinkGetLinkedCubinSize(h, &n);
    if (r != NVJITLINK_SUCCESS) return log_and_fail("nvJitLinkGetLinkedCubinSize", r);
    b200ode_rhs_s* H = new b200ode_rhs_s;
    H->solver = solver;
    H->neq = neq;
    H->n_events = n_events;
    H->flags = flags;
    H->cubin.resize(n);
    r = nvJitLinkGetLinkedCubin(h, H->cubin.data());
    if (r != NVJITLINK_SUCCESS) {
        delete H;
        return log_and_fail("nvJitLinkGetLinkedCubin", r);
    }
    nvJitLinkDestroy(&h);
    *out = H;
    return 0;
}

extern "C" int b200ode_link_rhs(int solver, int neq, const void* image, size_t size, int kind,
                                unsigned flags, b200ode_rhs_t* out)
{
    return b200ode_link_rhs_lanes(solver, neq, image, size, kind, nullptr, 0, 0, flags, out);
}

extern "C" int b200ode_link_rhs_lanes(int solver, int neq, const void* image, size_t size, int kind,
                                      const void* lanes_image, size_t lanes_size, int lanes_kind,
                                      unsigned flags, b200ode_rhs_t* out)
{
    if (!out) return fail(B200ODE_EINVAL, "out is NULL");
    *out = nullptr;
    if (solver == B200ODE_SOLVE_IVP)
        return b200ode_link_ivp(neq, image, size, kind, 0, nullptr, 0, 0, flags, out);
    if (solver != B200ODE_DOP853 && solver != B200ODE_LSODA)
        return fail(B200ODE_EINVAL, "unknown solver %d", solver);
    if (neq < 1) return fail(B200ODE_EINVAL, "neq = %d", neq);
    char name[64];
    if (warp_kernel(solver, neq)) {
        if (neq > B200ODE_WARP_NMAX)
            return fail(B200ODE_ENOSYS, "neq = %d exceeds %d", neq, B200ODE_WARP_NMAX);
        snprintf(name, sizeof name, solver == B200ODE_DOP853 ? "dop853w" : "lsodaw");
    } else {
        snprintf(name, sizeof name, "%s_n%d", solver == B200ODE_DOP853 ? "dop853" : "lsoda", neq);
    }
    UserImage rhs, lanes;
    int rc = resolve_image(image, size, kind, "rhs_", "right-hand side", &rhs);
    if (rc) return rc;
    if (lanes_image && (rc = resolve_image(lanes_image, lanes_size, lanes_kind, "rhs_", "lane-parallel right-hand side", &lanes)))
        return rc;
    return link_handle(solver, neq, name, rhs, nullptr, 0, flags, out, lanes_image ? &lanes : nullptr,
                       kind == B200ODE_IMG_BUILTIN);
}

extern "C" int b200ode_link_ivp(int neq, const void* rhs_image, size_t rhs_size, int rhs_kind,
                                int n_events, const void* events_image, size_t events_size,
                                int events_kind, unsigned flags, b200ode_rhs_t* out)
{
    if (!out) return fail(B200ODE_EINVAL, "out is NULL");
    *out = nullptr;
    if (neq < 1) return fail(B200ODE_EINVAL, "neq = %d", neq);
    if (neq > B200ODE_WARP_NMAX)
        return fail(B200ODE_ENOSYS, "solve_ivp: neq = %d exceeds %d", neq, B200ODE_WARP_NMAX);
    if (n_events < 0 || n_events > B200ODE_IVP_NEVMAX)
        return fail(n_events < 0 ? B200ODE_EINVAL : B200ODE_ENOSYS,
                    "solve_ivp: n_events = %d (at most %d)", n_events, B200ODE_IVP_NEVMAX);
    if (n_events > 0 && !events_image)
        return fail(B200ODE_EINVAL, "n_events = %d but the events image is NULL", n_events);
    char name[64];
    if (warp_kernel(B200ODE_SOLVE_IVP, neq)) snprintf(name, sizeof name, "ivpw");
    else snprintf(name, sizeof name, "ivp_n%d", neq);
    UserImage rhs, ev;
    int rc = resolve_image(rhs_image, rhs_size, rhs_kind, "rhs_", "right-hand side", &rhs);
    if (rc) return rc;
    if (n_events > 0)
        rc = resolve_image(events_image, events_size, events_kind, "events_", "events function", &ev);
    else
        rc = resolve_image("none", 4, B200ODE_IMG_BUILTIN, "events_", "events function", &ev);
    if (rc) return rc;
    return link_handle(B200ODE_SOLVE_IVP, neq, name, rhs, &ev, n_events, flags, out, nullptr,
                       rhs_kind == B200ODE_IMG_BUILTIN);
}

// Frees what init_device_state() created (the device's context must be current).
static void release_device_state(PerDevice& P)
{
    if (P.next) g_drv.cuMemFree_v2(P.next);
    if (P.wm_scratch) g_drv.cuMemFree_v2(P.wm_scratch);
    for (int i = 0; i < PerDevice::kScratchSets; i++)
        if (P.wm_done[i]) g_drv.cuEventDestroy_v2(P.wm_done[i]);
    if (P.mod) g_drv.cuModuleUnload(P.mod);
    P.next = 0, P.wm_scratch = 0, P.mod = nullptr, P.fn = P.fn_ptr = nullptr;
}

extern "C" void b200ode_rhs_free(b200ode_rhs_t H)
{
    if (!H) return;
    if (g_drv.ok) {
        CUcontext prev = nullptr;
        g_drv.cuCtxGetCurrent(&prev);
        for (auto& kv : H->dev) {
            if (!kv.second.ctx) continue;
            g_drv.cuCtxSetCurrent(kv.second.ctx);
            release_device_state(kv.second);
            CUdevice dev;
            if (g_drv.cuDeviceGet(&dev, kv.first) == CUDA_SUCCESS) g_drv.cuDevicePrimaryCtxRelease_v2(dev);
        }
        g_drv.cuCtxSetCurrent(prev);
    }
    delete H;
}

extern "C" int b200ode_rhs_cubin(b200ode_rhs_t H, const void** cubin, size_t* size)
{
    if (!H || !cubin || !size) return fail(B200ODE_EINVAL, "NULL argument");
    *cubin = H->cubin.data();
    *size = H->cubin.size();
    return 0;
}

// Loads the handle's cubin into the current context and fills `P` (module, functions, launch
// geometry, work-queue heads).  On failure the caller releases whatever was created.
static int init_device_state(b200ode_rhs_t H, int device, PerDevice& P)
{
    CUdevice dev;
    CU(cuDeviceGet(&dev, device));
    CU(cuModuleLoadData(&P.mod, H->cubin.data()));
    const bool warp = warp_kernel(H->solver, H->neq);
    CU(cuModuleGetFunction(&P.fn, P.mod, kernel_name(H->solver, H->neq)));
    if (warp) {
        P.fn_ptr = P.fn;  // the warp kernel always hands the data record over by address
        if (H->solver == B200ODE_LSODA)
            CU(cuModuleGetFunction(&P.fn_band, P.mod, "b200ode_lsodaw_band_kernel"));
        if (H->solver == B200ODE_LSODA && H->neq > 32 && H->neq <= B200ODE_LSODAW_TMEM_NMAX) {
            // (only if the eight warps' vectors fit the shared memory of a block)
            const int need = B200ODE_LSODAW_TMEM_WARPS * B200ODE_LSODAW_TMEM_DOUBLES(H->neq) * (int)sizeof(double);
            int optin = 0;
            CU(cuDeviceGetAttribute(&optin, CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN, dev));
            if (need + 2048 <= optin) {
                CU(cuModuleGetFunction(&P.fn_tmem, P.mod, "b200ode_lsodaw_tmem_kernel"));
                CU(cuFuncSetAttribute(P.fn_tmem, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, need));
            }
        }
    } else
        CU(cuModuleGetFunction(&P.fn_ptr, P.mod,
                               (std::string(kernel_name(H->solver, H->neq)) + "_ptr").c_str()));
    if (!warp && H->solver == B200ODE_DOP853) {
        CU(cuModuleGetFunction(&P.fn_vtol, P.mod, "b200ode_dop853_kernel_vtol"));
        CU(cuModuleGetFunction(&P.fn_ptr_vtol, P.mod, "b200ode_dop853_kernel_ptr_vtol"));
    }
    CU(cuMemAlloc_v2(&P.next, sizeof(unsigned long long) * kCounterSlots));
    CU(cuFuncGetAttribute(&P.regs, CU_FUNC_ATTRIBUTE_NUM_REGS, P.fn));
    CU(cuFuncGetAttribute(&P.local_bytes, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, P.fn));
    CU(cuFuncGetAttribute(&P.smem, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, P.fn));
    if (warp && H->solver == B200ODE_SOLVE_IVP) {
        P.threads = 32;
        P.dyn_smem = B200ODE_IVPW_DOUBLES(H->neq) * (int)sizeof(double);
    } else if (warp && H->solver == B200ODE_DOP853) {
        P.threads = 32;
        P.dyn_smem = B200ODE_DOP853W_DOUBLES(H->neq) * (int)sizeof(double);
    } else if (warp) {
        P.threads = 32;
        P.dyn_smem = B200ODE_LSODAW_DOUBLES(H->neq) * (int)sizeof(double);
        // B200ODE_WM=global: iteration matrices in global memory (more warps per SM;
        // measured slower at n = 64: 50 k vs 64 k traj/s, the LU then waits on L2)
        const char* wm = getenv("B200ODE_WM");
        if (wm && !strcmp(wm, "global"))
            P.dyn_smem -= B200ODE_LSODAW_MATRIX(H->neq) * (int)sizeof(double);
    } else {
        // ask the image for its build-time geometry (queue / storage sizes, threads)
        const bool dop = H->solver == B200ODE_DOP853;
        const bool ivp = H->solver == B200ODE_SOLVE_IVP;
        CUfunction cfg;
        CU(cuModuleGetFunction(&cfg, P.mod, ivp ? "b200ode_ivp_config"
                                                : dop ? "b200ode_dop853_config" : "b200ode_lsoda_config"));
        int* d_cfg = reinterpret_cast<int*>(P.next);  // scratch: 32 bytes of the slots
        void* cp[] = {&d_cfg};
        CU(cuLaunchKernel(cfg, 1, 1, 1, 1, 1, 1, 0, nullptr, cp, nullptr));
        int cfgv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        CU(cuMemcpyDtoH_v2(cfgv, P.next, sizeof cfgv));
        if (cfgv[3] != H->neq)
            return fail(B200ODE_ECUDA, "solver image was built for neq = %d", cfgv[3]);
        P.threads = cfgv[4];
        if (P.threads < 32 || P.threads > 1024 || P.threads % 32)
            return fail(B200ODE_ECUDA, "solver image reports %d threads per block", P.threads);
        if (ivp) {
            if (cfgv[0] != B200ODE_IVP_DOUBLES(H->neq) * P.threads || cfgv[1] != B200ODE_IVP_NEVMAX)
                return fail(B200ODE_ECUDA, "solve_ivp image reports an inconsistent geometry");
            P.dyn_smem = cfgv[0] * (int)sizeof(double);
        } else if (dop) {
            if (cfgv[0] < 32 || cfgv[1] != B200ODE_DOP853_QFIELDS(H->neq))
                return fail(B200ODE_ECUDA, "dop853 image reports an inconsistent queue geometry");
            P.qcap = cfgv[0];
            P.dyn_smem = cfgv[5] * (int)sizeof(double);
        } else {
            P.dyn_smem = cfgv[0] * (int)sizeof(double);
        }
    }
    P.max_dyn_smem = P.dyn_smem;
    if (!warp && H->solver == B200ODE_DOP853) {
        // room for a per-block copy of t_eval (launch(): te_shared)
        int optin = 0;
        CU(cuDeviceGetAttribute(&optin, CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN, dev));
        P.max_dyn_smem = std::max(P.dyn_smem, optin - P.smem);
    }
    if (P.max_dyn_smem > 48 * 1024) {
        CU(cuFuncSetAttribute(P.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, P.max_dyn_smem));
        if (P.fn_ptr != P.fn)
            CU(cuFuncSetAttribute(P.fn_ptr, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, P.max_dyn_smem));
        for (CUfunction f : {P.fn_band, P.fn_vtol, P.fn_ptr_vtol})
            if (f) CU(cuFuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, P.max_dyn_smem));
    }
    CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&P.blocks_per_sm, P.fn, P.threads, P.dyn_smem));
    CU(cuDeviceGetAttribute(&P.sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
    if (P.blocks_per_sm < 1) P.blocks_per_sm = 1;
    if (warp && H->solver == B200ODE_LSODA &&
        P.dyn_smem < B200ODE_LSODAW_DOUBLES(H->neq) * (int)sizeof(double)) {
        P.wm_set_bytes = (size_t)P.sms * P.blocks_per_sm * B200ODE_LSODAW_MATRIX(H->neq) * sizeof(double);
        CU(cuMemAlloc_v2(&P.wm_scratch, P.wm_set_bytes * PerDevice::kScratchSets));
        for (int i = 0; i < PerDevice::kScratchSets; i++)
            CU(cuEventCreate(&P.wm_done[i], CU_EVENT_DISABLE_TIMING));
    }
    return 0;
}

// Make `device`'s primary context current on this thread (shared with the CUDA runtime and
// torch) and load the module there once.  The per-device state is built in a local object
// and committed to the handle only when every step succeeded: a failed initialisation
// leaves nothing half-made behind, and the next call tries again.
static int device_state(b200ode_rhs_t H, int device, PerDevice** out)
{
    if (!driver())
        return fail(B200ODE_ENODEV, "CUDA driver (libcuda.so.1) not available: this library has "
                                    "no CPU fallback");
    CUdevice dev;
    if (device < 0) {
        CUcontext cur = nullptr;
        CU(cuCtxGetCurrent(&cur));
        if (!cur) {
            device = 0;
        } else {
            CU(cuCtxGetDevice(&dev));
            device = (int)dev;
        }
    }
    std::lock_guard<std::mutex> lk(H->mu);
    auto it = H->dev.find(device);
    if (it != H->dev.end()) {
        CU(cuCtxSetCurrent(it->second.ctx));
        *out = &it->second;
        return 0;
    }
    PerDevice P;
    CU(cuDeviceGet(&dev, device));
    CU(cuDevicePrimaryCtxRetain(&P.ctx, dev));
    int rc = 0;
    {
        CUresult r = g_drv.cuCtxSetCurrent(P.ctx);
        rc = r == CUDA_SUCCESS ? init_device_state(H, device, P) : cu_fail(r, "cuCtxSetCurrent");
    }
    if (rc) {
        const std::string msg = g_err;  // (the clean-up must not overwrite the message)
        release_device_state(P);
        g_drv.cuDevicePrimaryCtxRelease_v2(dev);
        g_err = msg;
        return rc;
    }
    PerDevice& Q = H->dev[device];
    Q = P;
    *out = &Q;
    return 0;
}

// Host-path calls name their device explicitly and make its context current for the duration
// of the call only: the caller's current context (torch's, the CUDA runtime's) is put back.
namespace {
struct CtxRestore {
    CUcontext prev = nullptr;
    bool armed = false;
    CtxRestore()
    {
        if (g_drv.ok && g_drv.cuCtxGetCurrent(&prev) == CUDA_SUCCESS) armed = true;
    }
    ~CtxRestore()
    {
        if (armed) g_drv.cuCtxSetCurrent(prev);
    }
};
}  // namespace

// A tolerance array of the C ABI: pointer + B200ODE_TOL_* shape -> what the kernels take
// (b200ode_launch.h: value of trajectory b, component i = p[b * sb + (vec ? i : 0)]).
struct TolArg {
    const double* p = nullptr;
    int dims = 0;
    long long sb(int neq) const { return dims == B200ODE_TOL_PER_COMPONENT ? 0 : dims == B200ODE_TOL_FULL ? neq : 1; }
    int vec() const { return p && dims != B200ODE_TOL_PER_TRAJECTORY ? 1 : 0; }
    size_t count(long long B, int neq) const
    {
        return dims == B200ODE_TOL_PER_COMPONENT ? (size_t)neq
                                                 : dims == B200ODE_TOL_FULL ? (size_t)B * neq : (size_t)B;
    }
};
static TolArg tol_arg(const double* p, int dims)
{
    TolArg t;
    t.p = p;
    t.dims = p ? dims : 0;
    return t;
}
static int check_tol(const TolArg& t, const char* what)
{
    if (t.dims < 0 || t.dims > B200ODE_TOL_FULL) return fail(B200ODE_EINVAL, "%s_dims = %d", what, t.dims);
    return 0;
}

// How lsoda forms its iteration matrix (b200ode_problem.jt / ml / mu): ml < 0 = dense.
struct JacArg {
    int jt = 0, ml = -1, mu = -1;
};
static JacArg jac_arg(int jt, int ml, int mu)
{
    JacArg j;
    j.jt = jt;
    if (jt == 5) j.ml = ml, j.mu = mu;
    return j;
}
static int check_jac(const JacArg& j, int solver, int neq)
{
    if (j.jt != 0 && j.jt != 2 && j.jt != 5)
        return fail(B200ODE_ENOSYS, "jt = %d: the Jacobian is generated internally, dense (jt = 2) or banded (jt = 5)", j.jt);
    if (j.jt == 5 && solver != B200ODE_LSODA) return fail(B200ODE_EINVAL, "jt = 5 is an lsoda option");
    // the reference's checker, LSODA.cpp:376-391
    if (j.jt == 5 && (j.ml < 0 || j.mu < 0 || j.ml >= neq || j.mu >= neq))
        return fail(B200ODE_EINVAL, "ml = %d, mu = %d not between 0 and neq - 1 = %d", j.ml, j.mu, neq - 1);
    return 0;
}

static int check_args(b200ode_rhs_t H, int solver, long long B, int neq, const void* u0,
                      const void* data, int np, int nt, const void* teval, const void* usol,
                      const void* success)
{
    if (!H) return fail(B200ODE_EINVAL, "rhs handle is NULL");
    if (H->solver != solver)
        return fail(B200ODE_EINVAL, "rhs handle was linked for the other solver");
    if (neq != H->neq)
        return fail(B200ODE_EINVAL, "neq = %d but the handle was linked for neq = %d", neq, H->neq);
    if (B < 0 || nt < 1) return fail(B200ODE_EINVAL, "B = %lld, nt = %d", B, nt);
    if (np > B200ODE_NPMAX)
        return fail(B200ODE_EINVAL, "np = %d exceeds %d; pass np = -1 to hand the record's "
                                    "address to the right-hand side instead", np, B200ODE_NPMAX);
    if (B > 0 && (!u0 || !teval || !usol || !success || (!data && np != 0)))
        return fail(B200ODE_EINVAL, "NULL array argument");
    return 0;
}

// Enqueue one ensemble on `stream`; all pointers are device pointers.
static int launch(b200ode_rhs_t H, PerDevice* P, const B200OdeLaunch& args_in, CUstream stream)
{
    if (args_in.B == 0) return 0;
    // (the thread kernels keep a lane's trajectory index in 32 bits; the host-buffer path launches
    // chunks far below this, a device-resident ensemble of more has to be split by the caller)
    if (args_in.B > 0x7fffffffLL)
        return fail(B200ODE_EINVAL, "B = %lld: at most 2^31 - 1 trajectories per launch", args_in.B);
    B200OdeLaunch args = args_in;
    unsigned slot = H->slot.fetch_add(1) % kCounterSlots;
    CUdeviceptr next = P->next + slot * sizeof(unsigned long long);
    CU(cuMemsetD8Async(next, 0, sizeof(unsigned long long), stream));
    args.next = reinterpret_cast<unsigned long long*>(next);
    const int kThreads = P->threads;
    const bool warp = warp_kernel(H->solver, H->neq);
    args.neq = H->neq;
    args.strict = (H->flags & B200ODE_STRICT_FP) ? 1 : 0;

    long long want = warp ? args.B : (args.B + kThreads - 1) / kThreads;
    long long resident = (long long)P->sms * P->blocks_per_sm;
    int grid = (int)std::min(want, resident);
    const int set = (int)(slot % PerDevice::kScratchSets);
    if (P->wm_scratch) {
        std::lock_guard<std::mutex> lk(H->mu);
        if (P->wm_used[set]) CU(cuStreamWaitEvent(stream, P->wm_done[set], 0));
        P->wm_used[set] = true;
        args.scratch = reinterpret_cast<double*>(P->wm_scratch + P->wm_set_bytes * set);
    }
    CUfunction fn = (args.np >= 0 && !warp) ? P->fn : P->fn_ptr;
    if (P->fn_vtol && (args.rtol_vec || args.atol_vec)) fn = args.np >= 0 ? P->fn_vtol : P->fn_ptr_vtol;
    int dyn_smem = P->dyn_smem;
    if (!(warp && H->solver == B200ODE_LSODA) || P->wm_scratch ||
        (args.ml >= 0 && B200ODE_LSODAW_BAND_DOUBLES(H->neq, args.ml, args.mu) >= B200ODE_LSODAW_DOUBLES(H->neq)))
        args.ml = args.mu = -1;  // thread kernels and bands as wide as the matrix: dense
    if (args.ml >= 0) {
        // banded iteration matrix: a trajectory needs far less shared memory, more warps fit an SM
        fn = P->fn_band;
        dyn_smem = B200ODE_LSODAW_BAND_DOUBLES(H->neq, args.ml, args.mu) * (int)sizeof(double);
        int nb = 0;
        CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, kThreads, (size_t)dyn_smem));
        if (nb < 1) nb = 1;
        grid = (int)std::min<long long>(want, (long long)P->sms * nb);
    }
    int threads = kThreads;
    if (P->fn_tmem && fn == P->fn_ptr && !P->wm_scratch && P->blocks_per_sm < B200ODE_LSODAW_TMEM_WARPS) {
        // dense matrix, fewer than eight trajectories per SM in shared memory: keep the matrices in
        // tensor memory instead (B200ODE_LSODAW_TMEM=0: the shared-memory kernel, for comparisons)
        const char* e = getenv("B200ODE_LSODAW_TMEM");
        if (!(e && !strcmp(e, "0"))) {
            fn = P->fn_tmem;
            threads = 32 * B200ODE_LSODAW_TMEM_WARPS;
            dyn_smem = B200ODE_LSODAW_TMEM_WARPS * B200ODE_LSODAW_TMEM_DOUBLES(H->neq) * (int)sizeof(double);
            grid = (int)std::min<long long>((want + B200ODE_LSODAW_TMEM_WARPS - 1) / B200ODE_LSODAW_TMEM_WARPS,
                                            (long long)P->sms);
        }
    }
    args.te_shared = 0;
    if (!warp && H->solver == B200ODE_DOP853) {
        // a copy of t_eval per block, if it costs no resident block (dop853_kernel.cu)
        const long long with_te = (long long)P->dyn_smem + (long long)args.nt * (long long)sizeof(double);
        int nb = 0;
        if (with_te <= P->max_dyn_smem &&
            g_drv.cuOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, kThreads, (size_t)with_te) == CUDA_SUCCESS &&
            nb >= P->blocks_per_sm) {
            args.te_shared = 1;
            dyn_smem = (int)with_te;
        }
    }
    void* params[] = {&args};
    CU(cuLaunchKernel(fn, grid, 1, 1, threads, 1, 1, (unsigned)dyn_smem, stream, params, nullptr));
    if (P->wm_scratch) CU(cuEventRecord(P->wm_done[set], stream));
    H->last_grid = grid;
    H->launches++;
    return 0;
}

static int solve_device(int solver, b200ode_rhs_t H, long long B, int neq, const double* u0,
                        long long u0_stride, const void* data, long long data_stride, int np,
                        int nt, const double* teval, double* usol, double rtol, double atol,
                        int mxstep, int* success, int* counters, int exit_on_warning,
                        void* stream, TolArg rt = TolArg(), TolArg at = TolArg(), JacArg jac = JacArg())
{
    int rc = check_args(H, solver, B, neq, u0, data, np, nt, teval, usol, success);
    if (rc) return rc;
    if ((rc = check_tol(rt, "rtol")) || (rc = check_tol(at, "atol")) || (rc = check_jac(jac, solver, neq))) return rc;
    if (u0_stride != 0 && u0_stride < neq) return fail(B200ODE_EINVAL, "u0_stride = %lld", u0_stride);
    PerDevice* P;
    rc = device_state(H, -1, &P);
    if (rc) return rc;
    B200OdeLaunch a;
    memset(&a, 0, sizeof a);
    a.B = B;
    a.u0 = u0;
    a.u0_stride = u0_stride;
    a.data = static_cast<const char*>(data);
    // data_stride < 0: one record of -data_stride bytes shared by all (include/b200ode.h);
    // the kernels address record b at data + b * stride, so shared means stride 0
    a.data_stride = data_stride > 0 ? data_stride : 0;
    a.np = np;
    a.nt = nt;
    a.t_eval = teval;
    a.usol = usol;
    a.rtol = rtol;
    a.atol = atol;
    a.rtol_b = rt.p;
    a.atol_b = at.p;
    a.rtol_sb = rt.sb(neq), a.rtol_vec = rt.vec();
    a.atol_sb = at.sb(neq), a.atol_vec = at.vec();
    a.mxstep = mxstep;
    a.exit_on_warning = exit_on_warning;
    a.success = success;
    a.counters = counters;
    a.ml = jac.ml, a.mu = jac.mu;
    return launch(H, P, a, static_cast<CUstream>(stream));
}

extern "C" int b200ode_dop853_batched_device(b200ode_rhs_t H, long long B, int neq,
                                             const double* u0, long long u0_stride,
                                             const void* data, long long data_stride, int np,
                                             int nt, const double* teval, double* usol,
                                             double rtol, double atol, int mxstep,
                                             int* success, int* counters, void* stream)
{
    return solve_device(B200ODE_DOP853, H, B, neq, u0, u0_stride, data, data_stride, np, nt,
                        teval, usol, rtol, atol, mxstep, success, counters, 0, stream);
}

extern "C" int b200ode_lsoda_batched_device(b200ode_rhs_t H, long long B, int neq,
                                            const double* u0, long long u0_stride,
                                            const void* data, long long data_stride, int np,
                                            int nt, const double* teval, double* usol,
                                            double rtol, double atol, int mxstep, int* success,
                                            int* counters, int exit_on_warning, void* stream)
{
    return solve_device(B200ODE_LSODA, H, B, neq, u0, u0_stride, data, data_stride, np, nt,
                        teval, usol, rtol, atol, mxstep, success, counters, exit_on_warning,
                        stream);
}

// ---------------------------------------------------------------- host-buffer path
namespace {
struct Stream {
    CUstream s = nullptr;
    ~Stream() { if (s && g_drv.ok) g_drv.cuStreamDestroy_v2(s); }
    int create() { CU(cuStreamCreate(&s, CU_STREAM_NON_BLOCKING)); return 0; }
};
struct Event {
    CUevent e = nullptr;
    ~Event() { if (e && g_drv.ok) g_drv.cuEventDestroy_v2(e); }
    int create() { CU(cuEventCreate(&e, CU_EVENT_DISABLE_TIMING)); return 0; }
};
}  // namespace

// Device scratch memory is kept between calls (per device, grow-only): cuMemAlloc /
// cuMemFree of gigabytes cost milliseconds each and serialise with running work.
namespace {
struct ScratchBlock {
    CUdeviceptr p = 0;
    size_t size = 0;
    bool busy = false;
};
struct ScratchPool {
    std::mutex mu;
    std::map<int, std::vector<ScratchBlock>> blocks;  // by device
};
ScratchPool g_pool;

// One allocation carved into the buffers of a call; returned to the pool on scope exit.
struct Scratch {
    int device = -1;
    size_t index = 0;
    CUdeviceptr base = 0;
    size_t used = 0, size = 0;
    bool pooled = false;
    int acquire(int dev, size_t bytes)
    {
        device = dev;
        std::lock_guard<std::mutex> lk(g_pool.mu);
        auto& v = g_pool.blocks[dev];
        for (size_t i = 0; i < v.size(); i++)
            if (!v[i].busy && v[i].size >= bytes) {
                v[i].busy = true;
                index = i, base = v[i].p, size = v[i].size, pooled = true;
                return 0;
            }
        for (size_t i = 0; i < v.size(); i++)
            if (!v[i].busy) {  // too small: replace it
                g_drv.cuMemFree_v2(v[i].p);
                v[i].p = 0, v[i].size = 0;
                CU(cuMemAlloc_v2(&v[i].p, bytes));
                v[i].size = bytes, v[i].busy = true;
                index = i, base = v[i].p, size = bytes, pooled = true;
                return 0;
            }
        ScratchBlock b;
        CU(cuMemAlloc_v2(&b.p, bytes));
        b.size = bytes, b.busy = true;
        v.push_back(b);
        index = v.size() - 1, base = b.p, size = bytes, pooled = true;
        return 0;
    }
    CUdeviceptr take(size_t bytes)
    {
        CUdeviceptr p = base + used;
        used += (bytes + 255) & ~(size_t)255;
        return p;
    }
    // work that uses the block could not be waited for: never hand it out again (it stays
    // marked busy -- leaked on purpose, freeing it could fault the work still running)
    bool poisoned = false;
    ~Scratch()
    {
        if (!pooled || poisoned) return;
        std::lock_guard<std::mutex> lk(g_pool.mu);
        g_pool.blocks[device][index].busy = false;
    }
};

// Waits for the streams of a host-path call on every way out of it.  Once the first copy or
// kernel is enqueued, an early error return must not release the pooled device block (the
// next call would reuse it under running kernels) nor give the caller's host buffers back
// while asynchronous copies may still write into them.
struct StreamDrain {
    Scratch* mem;
    CUstream s[3] = {nullptr, nullptr, nullptr};
    explicit StreamDrain(Scratch* m) : mem(m) {}
    int wait()
    {
        int rc = 0;
        for (CUstream q : s)
            if (q) {
                CUresult r = g_drv.cuStreamSynchronize(q);
                if (r != CUDA_SUCCESS) {
                    if (mem) mem->poisoned = true;
                    if (!rc) rc = cu_fail(r, "cuStreamSynchronize");
                }
            }
        s[0] = s[1] = s[2] = nullptr;
        return rc;
    }
    ~StreamDrain()
    {
        const std::string msg = g_err;  // keep the error that caused the early return
        wait();
        if (!msg.empty()) g_err = msg;
    }
};
static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }
static double now_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
}  // namespace

// Host-buffer entry point.  The ensemble is cut into chunks; chunk c integrates on one
// of two kernel streams while the rows of chunk c-1 travel back over PCIe on the copy
// stream and chunk c+1's blocks move onto SMs as chunk c's blocks retire (each chunk is
// a persistent grid, so the tail of one launch is filled by the head of the next).
// Output buffers form a ring of three (kernel c, kernel c+1, copy c-1).
static int solve_host(int solver, b200ode_rhs_t H, long long B, int neq, const double* u0,
                      long long u0_stride, const void* data, long long data_stride, int np,
                      int nt, const double* teval, double* usol, double rtol, double atol,
                      int mxstep, int* success, int* counters, int exit_on_warning, int device,
                      TolArg rt = TolArg(), TolArg at = TolArg(), JacArg jac = JacArg())
{
    int rc = check_args(H, solver, B, neq, u0, data, np, nt, teval, usol, success);
    if (rc) return rc;
    if ((rc = check_tol(rt, "rtol")) || (rc = check_tol(at, "atol")) || (rc = check_jac(jac, solver, neq))) return rc;
    const double* rtol_b = rt.p;
    const double* atol_b = at.p;
    if (device < 0) return fail(B200ODE_EINVAL, "device = %d", device);
    if (u0_stride != 0 && u0_stride < neq) return fail(B200ODE_EINVAL, "u0_stride = %lld", u0_stride);
    if (B == 0) return 0;
    // dop853: nmax <= 0 is rejected before any work (dop853_module.f90:243-246)
    if (solver == B200ODE_DOP853 && mxstep <= 0) {
        for (long long b = 0; b < B; b++) success[b] = 0;
        if (counters) {
            memset(counters, 0, sizeof(int) * B200ODE_NCOUNTERS * (size_t)B);
            for (long long b = 0; b < B; b++) counters[b * B200ODE_NCOUNTERS + 5] = -1;
        }
        return 0;
    }
    const bool trace = getenv("B200ODE_TRACE") != nullptr;
    const double t0 = now_ms();
    CtxRestore restore_ctx;
    PerDevice* P;
    rc = device_state(H, device, &P);
    if (rc) return rc;

    const size_t row_bytes = sizeof(double) * (size_t)nt * (size_t)neq;
    const size_t per_traj = row_bytes + sizeof(int) + (counters ? sizeof(int) * B200ODE_NCOUNTERS : 0);
    // record size: |data_stride| bytes, or np doubles when data_stride == 0
    const size_t rec = data_stride != 0 ? (size_t)(data_stride > 0 ? data_stride : -data_stride)
                                        : (size_t)(np > 0 ? np * 8 : 0);
    const size_t data_bytes = data ? (data_stride > 0 ? (size_t)data_stride * (size_t)B : rec) : 0;
    const size_t rtol_bytes = sizeof(double) * rt.count(B, neq), atol_bytes = sizeof(double) * at.count(B, neq);
    // u0 rows wider than neq travel as the padded block they are (one copy; the kernels take
    // the stride), not as B small copies
    const size_t u0_doubles = u0_stride ? (size_t)(B - 1) * (size_t)u0_stride + (size_t)neq : (size_t)neq;
    const size_t in_bytes = align256(sizeof(double) * nt) +
                            align256(sizeof(double) * u0_doubles) +
                            align256(data_bytes) + (rtol_b ? align256(rtol_bytes) : 0) +
                            (atol_b ? align256(atol_bytes) : 0);
    // chunk size.  Compute-bound shapes (few output bytes per trajectory): one trajectory per
    // resident lane and at least 1/16 of the ensemble -- the chunks run two at a time on the two
    // kernel streams, so a chunk's stragglers share the SMs with the next chunk's head, and what
    // stays exposed is the last chunk's copy (Robertson, 2^20 trajectories, e2e: 4 per lane
    // 8.05 M traj/s, 2 per lane 8.19 M, 1 per lane 8.51 M, 1/2 per lane 7.19 M).  Copy-bound shapes (>= 8 KB of rows per trajectory: the
    // device-to-host copy of a chunk takes longer than integrating it, e.g. Lorenz with 1001
    // rows): small chunks, half the resident lanes each -- two chunks run side by side on the
    // two kernel streams -- so that the first rows are on their way after one trajectory's
    // latency and only that first kernel is exposed.  At most 1 GiB of rows per buffer.
    const long long resident = (long long)P->sms * P->blocks_per_sm * (warp_kernel(solver, neq) ? 1 : P->threads);
    size_t free_b = 0, total_b = 0;
    CU(cuMemGetInfo_v2(&free_b, &total_b));
    const bool copy_bound = row_bytes >= 8192;
    long long chunk = copy_bound ? std::max<long long>(resident / 2, (B + 15) / 16)
                                 : std::max<long long>(resident, (B + 15) / 16);
    if (const char* e = getenv("B200ODE_CHUNK")) chunk = std::max<long long>(1, atoll(e));  // experiments
    const size_t cap = std::min<size_t>((size_t)1 << 30, std::max<size_t>(free_b / 8, (size_t)64 << 20));
    chunk = std::min<long long>(chunk, (long long)std::max<size_t>(1, cap / per_traj));
    chunk = std::max<long long>(1, std::min(chunk, B));
    const long long nchunks = (B + chunk - 1) / chunk;
    const int nbuf = (int)std::min<long long>(nchunks, 3);

    Scratch mem;
    const size_t out_bytes = align256(row_bytes * (size_t)chunk) + align256(sizeof(int) * (size_t)chunk) +
                             (counters ? align256(sizeof(int) * B200ODE_NCOUNTERS * (size_t)chunk) : 0);
    if ((rc = mem.acquire(device, in_bytes + out_bytes * nbuf + 4096))) return rc;
    const CUdeviceptr d_te = mem.take(sizeof(double) * nt);
    const CUdeviceptr d_u0 = mem.take(sizeof(double) * u0_doubles);
    const CUdeviceptr d_data = mem.take(data_bytes);
    const CUdeviceptr d_rtol = rtol_b ? mem.take(rtol_bytes) : 0;
    const CUdeviceptr d_atol = atol_b ? mem.take(atol_bytes) : 0;
    CUdeviceptr d_usol[3], d_ok[3], d_cnt[3] = {0, 0, 0};
    for (int i = 0; i < nbuf; i++) {
        d_usol[i] = mem.take(row_bytes * (size_t)chunk);
        d_ok[i] = mem.take(sizeof(int) * (size_t)chunk);
        if (counters) d_cnt[i] = mem.take(sizeof(int) * B200ODE_NCOUNTERS * (size_t)chunk);
    }
    Stream sk[2], sc;
    if ((rc = sk[0].create()) || (rc = sk[1].create()) || (rc = sc.create())) return rc;
    Event ev_in, ev_kernel[3], ev_copied[3];
    if ((rc = ev_in.create())) return rc;
    for (int i = 0; i < nbuf; i++)
        if ((rc = ev_kernel[i].create()) || (rc = ev_copied[i].create())) return rc;
    const double t1 = now_ms();
    StreamDrain drain(&mem);  // from here on every return waits for the streams first
    drain.s[0] = sk[0].s, drain.s[1] = sk[1].s, drain.s[2] = sc.s;

    CU(cuMemcpyHtoDAsync_v2(d_te, teval, sizeof(double) * nt, sk[0].s));
    CU(cuMemcpyHtoDAsync_v2(d_u0, u0, sizeof(double) * u0_doubles, sk[0].s));
    if (data_bytes) CU(cuMemcpyHtoDAsync_v2(d_data, data, data_bytes, sk[0].s));
    if (rtol_b) CU(cuMemcpyHtoDAsync_v2(d_rtol, rtol_b, rtol_bytes, sk[0].s));
    if (atol_b) CU(cuMemcpyHtoDAsync_v2(d_atol, atol_b, atol_bytes, sk[0].s));
    CU(cuEventRecord(ev_in.e, sk[0].s));
    CU(cuStreamWaitEvent(sk[1].s, ev_in.e, 0));

    auto copy_back = [&](long long c) -> int {
        const int i = (int)(c % nbuf);
        const long long b0 = c * chunk, nb = std::min(chunk, B - b0);
        CU(cuStreamWaitEvent(sc.s, ev_kernel[i].e, 0));
        CU(cuMemcpyDtoHAsync_v2(usol + (size_t)b0 * nt * neq, d_usol[i], row_bytes * (size_t)nb, sc.s));
        CU(cuMemcpyDtoHAsync_v2(success + b0, d_ok[i], sizeof(int) * (size_t)nb, sc.s));
        if (counters)
            CU(cuMemcpyDtoHAsync_v2(counters + b0 * B200ODE_NCOUNTERS, d_cnt[i],
                                    sizeof(int) * B200ODE_NCOUNTERS * (size_t)nb, sc.s));
        CU(cuEventRecord(ev_copied[i].e, sc.s));
        return 0;
    };
    for (long long c = 0; c < nchunks; c++) {
        const int i = (int)(c % nbuf);
        CUstream ks = sk[c & 1].s;
        const long long b0 = c * chunk, nb = std::min(chunk, B - b0);
        if (c >= nbuf) CU(cuStreamWaitEvent(ks, ev_copied[i].e, 0));  // buffer i is free again
        B200OdeLaunch a;
        memset(&a, 0, sizeof a);
        a.B = nb;
        a.u0 = reinterpret_cast<const double*>(d_u0) + (size_t)b0 * (size_t)u0_stride;
        a.u0_stride = u0_stride;
        a.data = reinterpret_cast<const char*>(d_data) + (data_stride > 0 ? (size_t)b0 * data_stride : 0);
        a.data_stride = data_stride > 0 ? data_stride : 0;
        a.np = np;
        a.nt = nt;
        a.t_eval = reinterpret_cast<const double*>(d_te);
        a.usol = reinterpret_cast<double*>(d_usol[i]);
        a.rtol = rtol;
        a.atol = atol;
        a.rtol_b = rtol_b ? reinterpret_cast<const double*>(d_rtol) + b0 * rt.sb(neq) : nullptr;
        a.atol_b = atol_b ? reinterpret_cast<const double*>(d_atol) + b0 * at.sb(neq) : nullptr;
        a.rtol_sb = rt.sb(neq), a.rtol_vec = rt.vec();
        a.atol_sb = at.sb(neq), a.atol_vec = at.vec();
        a.mxstep = mxstep;
        a.exit_on_warning = exit_on_warning;
        a.success = reinterpret_cast<int*>(d_ok[i]);
        a.counters = counters ? reinterpret_cast<int*>(d_cnt[i]) : nullptr;
        a.ml = jac.ml, a.mu = jac.mu;
        if ((rc = launch(H, P, a, ks))) return rc;
        CU(cuEventRecord(ev_kernel[i].e, ks));
        // the copy of chunk c-1 is enqueued after the launch of chunk c so that a buffer's
        // "copied" event always exists before a later chunk waits on it
        if (c >= 1 && (rc = copy_back(c - 1))) return rc;
    }
    if ((rc = copy_back(nchunks - 1))) return rc;
    const double t2 = now_ms();
    if ((rc = drain.wait())) return rc;
    if (trace)
        fprintf(stderr, "[b200ode] B=%lld chunk=%lld x%lld buffers=%d scratch=%.1f MB | setup %.2f ms, "
                        "enqueue %.2f ms, drain %.2f ms\n", B, chunk, nchunks, nbuf,
                (in_bytes + out_bytes * nbuf) / 1048576.0, t1 - t0, t2 - t1, now_ms() - t2);
    return 0;
}

extern "C" int b200ode_dop853_batched(b200ode_rhs_t H, long long B, int neq, const double* u0,
                                      long long u0_stride, const void* data,
                                      long long data_stride, int np, int nt,
                                      const double* teval, double* usol, double rtol,
                                      double atol, int mxstep, int* success, int* counters,
                                      int device)
{
    return solve_host(B200ODE_DOP853, H, B, neq, u0, u0_stride, data, data_stride, np, nt,
                      teval, usol, rtol, atol, mxstep, success, counters, 0, device);
}

extern "C" int b200ode_lsoda_batched(b200ode_rhs_t H, long long B, int neq, const double* u0,
                                     long long u0_stride, const void* data,
                                     long long data_stride, int np, int nt, const double* teval,
                                     double* usol, double rtol, double atol, int mxstep,
                                     int* success, int* counters, int exit_on_warning,
                                     int device)
{
    return solve_host(B200ODE_LSODA, H, B, neq, u0, u0_stride, data, data_stride, np, nt, teval,
                      usol, rtol, atol, mxstep, success, counters, exit_on_warning, device);
}

// ---------------------------------------------------------------- structure-based entry points
static int check_problem(b200ode_rhs_t H, const b200ode_problem* p)
{
    if (!H) return fail(B200ODE_EINVAL, "rhs handle is NULL");
    if (!p) return fail(B200ODE_EINVAL, "problem is NULL");
    if (p->size != sizeof(b200ode_problem))
        return fail(B200ODE_EINVAL, "problem.size = %zu, expected sizeof(b200ode_problem) = %zu",
                    p->size, sizeof(b200ode_problem));
    return 0;
}

extern "C" int b200ode_solve(b200ode_rhs_t H, const b200ode_problem* p, int device)
{
    int rc = check_problem(H, p);
    if (rc) return rc;
    return solve_host(H->solver, H, p->B, p->neq, p->u0, p->u0_stride, p->data, p->data_stride,
                      p->np, p->nt, p->teval, p->usol, p->rtol, p->atol, p->mxstep, p->success,
                      p->counters, H->solver == B200ODE_LSODA ? p->exit_on_warning : 0, device,
                      tol_arg(p->rtol_b, p->rtol_dims), tol_arg(p->atol_b, p->atol_dims),
                      jac_arg(p->jt, p->ml, p->mu));
}

extern "C" int b200ode_solve_device(b200ode_rhs_t H, const b200ode_problem* p, void* stream)
{
    int rc = check_problem(H, p);
    if (rc) return rc;
    return solve_device(H->solver, H, p->B, p->neq, p->u0, p->u0_stride, p->data,
                        p->data_stride, p->np, p->nt, p->teval, p->usol, p->rtol, p->atol,
                        p->mxstep, p->success, p->counters,
                        H->solver == B200ODE_LSODA ? p->exit_on_warning : 0, stream,
                        tol_arg(p->rtol_b, p->rtol_dims), tol_arg(p->atol_b, p->atol_dims),
                      jac_arg(p->jt, p->ml, p->mu));
}

// ---------------------------------------------------------------- solve_ivp
static const char* const kIvpMessages[] = {
    // the reference's strings, src/dop853_solve_ivp.f90:390, :364, :351, :300, :273
    "The solver successfully reached the end of the integration interval.",
    "A termination event occurred.",
    "Failed to complete integration.",
    "\"t_eval\" does not always increase or decrease",
    "Integrator failed to initialize.",
    "brent failed",              // the reference prints this and stops the process (:110-113)
    "output capacity exceeded",  // not in the reference, whose result arrays grow
};
extern "C" const char* b200ode_ivp_message(int code)
{
    if (code < 0 || code >= (int)(sizeof kIvpMessages / sizeof kIvpMessages[0])) return "";
    return kIvpMessages[code];
}

static int check_ivp(b200ode_rhs_t H, const b200ode_ivp_problem* p)
{
    if (!H) return fail(B200ODE_EINVAL, "handle is NULL");
    if (!p) return fail(B200ODE_EINVAL, "problem is NULL");
    if (p->size != sizeof(b200ode_ivp_problem))
        return fail(B200ODE_EINVAL, "problem.size = %zu, expected sizeof(b200ode_ivp_problem) = %zu",
                    p->size, sizeof(b200ode_ivp_problem));
    if (H->solver != B200ODE_SOLVE_IVP)
        return fail(B200ODE_EINVAL, "handle was not linked by b200ode_link_ivp");
    if (p->neq != H->neq)
        return fail(B200ODE_EINVAL, "neq = %d but the handle was linked for neq = %d", p->neq, H->neq);
    if (p->n_events != H->n_events)
        return fail(B200ODE_EINVAL, "n_events = %d but the handle was linked for %d", p->n_events,
                    H->n_events);
    if (p->B < 0 || p->nt < 0 || p->cap < 0)
        return fail(B200ODE_EINVAL, "B = %lld, nt = %d, cap = %lld", p->B, p->nt, p->cap);
    if (p->np > B200ODE_NPMAX)
        return fail(B200ODE_EINVAL, "np = %d exceeds %d; pass np = -1 to hand the record's address "
                                    "to the right-hand side instead", p->np, B200ODE_NPMAX);
    if (p->t_span_stride != 0 && p->t_span_stride != 2)
        return fail(B200ODE_EINVAL, "t_span_stride = %lld (0 or 2)", p->t_span_stride);
    if (p->y0_stride != 0 && p->y0_stride < p->neq)
        return fail(B200ODE_EINVAL, "y0_stride = %lld", p->y0_stride);
    if (p->B == 0) return 0;
    if (!p->t_span || !p->y0 || !p->result || !p->t_event || !p->y_event || (!p->data && p->np != 0) ||
        (p->nt > 0 && (!p->teval || !p->y_out)))
        return fail(B200ODE_EINVAL, "NULL array argument");
    return 0;
}

static int launch_ivp(b200ode_rhs_t H, PerDevice* P, B200IvpLaunch args, CUstream stream)
{
    if (args.B > 0x7fffffffLL)
        return fail(B200ODE_EINVAL, "B = %lld: at most 2^31 - 1 trajectories per launch", args.B);
    if (args.B == 0) return 0;
    unsigned slot = H->slot.fetch_add(1) % kCounterSlots;
    CUdeviceptr next = P->next + slot * sizeof(unsigned long long);
    CU(cuMemsetD8Async(next, 0, sizeof(unsigned long long), stream));
    args.next = reinterpret_cast<unsigned long long*>(next);
    const bool warp = warp_kernel(H->solver, H->neq);
    const long long want = warp ? args.B : (args.B + P->threads - 1) / P->threads;
    const long long resident = (long long)P->sms * P->blocks_per_sm;
    const int grid = (int)std::min(want, resident);
    CUfunction fn = (args.np >= 0 && !warp) ? P->fn : P->fn_ptr;
    int neq = H->neq;
    void* params[] = {&args, &neq};  // (the thread kernels take the first only)
    CU(cuLaunchKernel(fn, grid, 1, 1, P->threads, 1, 1, (unsigned)P->dyn_smem, stream, params, nullptr));
    H->last_grid = grid;
    H->launches++;
    return 0;
}

static void fill_ivp_launch(B200IvpLaunch* a, const b200ode_ivp_problem* p)
{
    memset(a, 0, sizeof *a);
    a->B = p->B;
    a->t_span = p->t_span;
    a->t_span_stride = p->t_span_stride;
    a->y0 = p->y0;
    a->y0_stride = p->y0_stride;
    a->data = static_cast<const char*>(p->data);
    a->data_stride = p->data_stride > 0 ? p->data_stride : 0;
    a->np = p->np;
    a->nt = p->nt;
    a->t_eval = p->teval;
    a->n_events = p->n_events;
    a->cap = p->nt > 0 ? p->nt : p->cap;
    a->row_offset = p->nt > 0 ? nullptr : p->row_offset;
    a->first_step = p->first_step;
    a->max_step = p->max_step;
    a->rtol = p->rtol;
    a->atol = p->atol;
    a->rtol_b = p->rtol_b;
    a->atol_b = p->atol_b;
    a->t_out = p->t_out;
    a->y_out = p->y_out;
    a->result = p->result;
    a->t_event = p->t_event;
    a->y_event = p->y_event;
}

extern "C" int b200ode_solve_ivp_device(b200ode_rhs_t H, const b200ode_ivp_problem* p, void* stream)
{
    int rc = check_ivp(H, p);
    if (rc) return rc;
    PerDevice* P;
    rc = device_state(H, -1, &P);
    if (rc) return rc;
    B200IvpLaunch a;
    fill_ivp_launch(&a, p);
    return launch_ivp(H, P, a, static_cast<CUstream>(stream));
}

// Host-buffer entry point: the ensemble is cut into chunks of trajectories whose rows fit a
// device buffer; chunk c integrates on the kernel stream while the rows of chunk c-1 travel
// back on the copy stream (two output buffers).
extern "C" int b200ode_solve_ivp(b200ode_rhs_t H, const b200ode_ivp_problem* p, int device)
{
    int rc = check_ivp(H, p);
    if (rc) return rc;
    if (device < 0) return fail(B200ODE_EINVAL, "device = %d", device);
    const long long B = p->B;
    if (B == 0) return 0;
    const int n = p->neq, nt = p->nt;
    const bool ragged = nt == 0 && p->row_offset != nullptr;
    if (ragged)
        for (long long b = 0; b < B; b++)
            if (p->row_offset[b + 1] < p->row_offset[b] || p->row_offset[b + 1] - p->row_offset[b] > 0x7fffffffLL)
                return fail(B200ODE_EINVAL, "row_offset must not decrease (trajectory %lld)", b);
    // first row of trajectory b in t_out / y_out (ragged: row_offset is used as given, so a
    // shard of a larger problem passes a slice of the offsets and the unshifted arrays)
    auto rows_upto = [&](long long b) -> long long {
        return nt > 0 ? b * nt : ragged ? p->row_offset[b] : b * p->cap;
    };
    if (rows_upto(B) > rows_upto(0) && (!p->y_out || (nt == 0 && !p->t_out)))
        return fail(B200ODE_EINVAL, "NULL output array");
    CtxRestore restore_ctx;
    PerDevice* P;
    rc = device_state(H, device, &P);
    if (rc) return rc;

    // chunks: at most ~4 GiB of rows (and an eighth of what the device has free), at least one trajectory
    size_t free_b = 0, total_b = 0;
    CU(cuMemGetInfo_v2(&free_b, &total_b));
    const size_t row_bytes = sizeof(double) * (size_t)(n + (nt == 0 ? 1 : 0));
    // (4 GiB: without t_eval a trajectory owns hundreds of rows, and a chunk should still hold
    // a few trajectories per resident lane)
    const size_t budget = std::min<size_t>((size_t)4 << 30, std::max<size_t>(free_b / 8, (size_t)64 << 20));
    const size_t fixed = sizeof(int) * B200ODE_IVP_NRESULT + sizeof(double) * (size_t)(1 + n);
    std::vector<long long> cut = {0};
    {
        long long b0 = 0;
        while (b0 < B) {
            long long b1 = b0 + 1;
            // grow the chunk while it fits the budget (binary search on the prefix sums)
            long long lo = b0 + 1, hi = B;
            while (lo < hi) {
                const long long mid = lo + (hi - lo + 1) / 2;
                const size_t bytes = (size_t)(rows_upto(mid) - rows_upto(b0)) * row_bytes + (size_t)(mid - b0) * fixed;
                if (bytes <= budget) lo = mid;
                else hi = mid - 1;
            }
            b1 = std::max(b1, lo);
            // a few chunks even when everything fits, so that copies overlap kernels
            const long long resident = (long long)P->sms * P->blocks_per_sm *
                                       (warp_kernel(H->solver, H->neq) ? 1 : P->threads);
            long long want = std::max<long long>(4 * resident, (B + 7) / 8);
            if (const char* e = getenv("B200ODE_IVP_CHUNK"))  // tests: force many small chunks
                want = std::max<long long>(1, atoll(e));
            b1 = std::min(b1, b0 + want);
            cut.push_back(b1);
            b0 = b1;
        }
    }
    const long long nchunks = (long long)cut.size() - 1;
    long long max_rows = 0, max_traj = 0;
    for (long long c = 0; c < nchunks; c++) {
        max_rows = std::max(max_rows, rows_upto(cut[c + 1]) - rows_upto(cut[c]));
        max_traj = std::max(max_traj, cut[c + 1] - cut[c]);
    }
    const int nbuf = (int)std::min<long long>(nchunks, 2);

    const size_t rec = p->data_stride != 0 ? (size_t)(p->data_stride > 0 ? p->data_stride : -p->data_stride)
                                           : (size_t)(p->np > 0 ? p->np * 8 : 0);
    const size_t data_bytes = p->data ? (p->data_stride > 0 ? (size_t)p->data_stride * (size_t)B : rec) : 0;
    const size_t ts_bytes = sizeof(double) * 2 * (size_t)(p->t_span_stride ? B : 1);
    // (rows wider than neq travel as the padded block they are: one copy, the kernel takes the stride)
    const size_t y0_bytes = sizeof(double) * (p->y0_stride ? (size_t)(B - 1) * (size_t)p->y0_stride + (size_t)n
                                                           : (size_t)n);
    const size_t tol_bytes = sizeof(double) * (size_t)B;
    const size_t off_bytes = ragged ? sizeof(long long) * (size_t)(max_traj + 1) : 0;
    const size_t in_bytes = align256(sizeof(double) * std::max(nt, 1)) + align256(ts_bytes) + align256(y0_bytes) +
                            align256(data_bytes) + (p->rtol_b ? align256(tol_bytes) : 0) +
                            (p->atol_b ? align256(tol_bytes) : 0);
    const size_t out_bytes = align256(sizeof(double) * (size_t)max_rows * n) +
                             (nt == 0 ? align256(sizeof(double) * (size_t)max_rows) : 0) +
                             align256(sizeof(int) * B200ODE_IVP_NRESULT * (size_t)max_traj) +
                             align256(sizeof(double) * (size_t)max_traj) +
                             align256(sizeof(double) * n * (size_t)max_traj) + align256(off_bytes);
    Scratch mem;
    if ((rc = mem.acquire(device, in_bytes + out_bytes * nbuf + 8192))) return rc;
    const CUdeviceptr d_te = mem.take(sizeof(double) * std::max(nt, 1));
    const CUdeviceptr d_ts = mem.take(ts_bytes);
    const CUdeviceptr d_y0 = mem.take(y0_bytes);
    const CUdeviceptr d_data = mem.take(data_bytes);
    const CUdeviceptr d_rtol = p->rtol_b ? mem.take(tol_bytes) : 0;
    const CUdeviceptr d_atol = p->atol_b ? mem.take(tol_bytes) : 0;
    CUdeviceptr d_y[2], d_t[2] = {0, 0}, d_res[2], d_tev[2], d_yev[2], d_off[2] = {0, 0};
    for (int i = 0; i < nbuf; i++) {
        d_y[i] = mem.take(sizeof(double) * (size_t)max_rows * n);
        if (nt == 0) d_t[i] = mem.take(sizeof(double) * (size_t)max_rows);
        d_res[i] = mem.take(sizeof(int) * B200ODE_IVP_NRESULT * (size_t)max_traj);
        d_tev[i] = mem.take(sizeof(double) * (size_t)max_traj);
        d_yev[i] = mem.take(sizeof(double) * n * (size_t)max_traj);
        if (ragged) d_off[i] = mem.take(off_bytes);
    }
    Stream sk, sc;
    if ((rc = sk.create()) || (rc = sc.create())) return rc;
    Event ev_kernel[2], ev_copied[2];
    for (int i = 0; i < nbuf; i++)
        if ((rc = ev_kernel[i].create()) || (rc = ev_copied[i].create())) return rc;

    StreamDrain drain(&mem);  // from here on every return waits for the streams first
    drain.s[0] = sk.s, drain.s[1] = sc.s;
    if (nt > 0) CU(cuMemcpyHtoDAsync_v2(d_te, p->teval, sizeof(double) * nt, sk.s));
    CU(cuMemcpyHtoDAsync_v2(d_ts, p->t_span, ts_bytes, sk.s));
    CU(cuMemcpyHtoDAsync_v2(d_y0, p->y0, y0_bytes, sk.s));
    if (data_bytes) CU(cuMemcpyHtoDAsync_v2(d_data, p->data, data_bytes, sk.s));
    if (p->rtol_b) CU(cuMemcpyHtoDAsync_v2(d_rtol, p->rtol_b, tol_bytes, sk.s));
    if (p->atol_b) CU(cuMemcpyHtoDAsync_v2(d_atol, p->atol_b, tol_bytes, sk.s));

    // ragged offsets of a chunk, relative to its first row (kept alive until the end)
    std::vector<std::vector<long long>> rel(ragged ? (size_t)nchunks : 0);
    auto copy_back = [&](long long c) -> int {
        const int i = (int)(c % nbuf);
        const long long b0 = cut[c], nb = cut[c + 1] - b0;
        const long long r0 = rows_upto(b0), nr = rows_upto(cut[c + 1]) - r0;
        CU(cuStreamWaitEvent(sc.s, ev_kernel[i].e, 0));
        if (nr > 0) {
            CU(cuMemcpyDtoHAsync_v2(p->y_out + (size_t)r0 * n, d_y[i], sizeof(double) * (size_t)nr * n, sc.s));
            if (nt == 0) CU(cuMemcpyDtoHAsync_v2(p->t_out + r0, d_t[i], sizeof(double) * (size_t)nr, sc.s));
        }
        CU(cuMemcpyDtoHAsync_v2(p->result + b0 * B200ODE_IVP_NRESULT, d_res[i],
                                sizeof(int) * B200ODE_IVP_NRESULT * (size_t)nb, sc.s));
        CU(cuMemcpyDtoHAsync_v2(p->t_event + b0, d_tev[i], sizeof(double) * (size_t)nb, sc.s));
        CU(cuMemcpyDtoHAsync_v2(p->y_event + b0 * n, d_yev[i], sizeof(double) * n * (size_t)nb, sc.s));
        CU(cuEventRecord(ev_copied[i].e, sc.s));
        return 0;
    };
    for (long long c = 0; c < nchunks; c++) {
        const int i = (int)(c % nbuf);
        const long long b0 = cut[c], nb = cut[c + 1] - b0;
        if (c >= nbuf) CU(cuStreamWaitEvent(sk.s, ev_copied[i].e, 0));  // buffer i is free again
        B200IvpLaunch a;
        fill_ivp_launch(&a, p);
        a.B = nb;
        a.t_span = reinterpret_cast<const double*>(d_ts) + (p->t_span_stride ? 2 * (size_t)b0 : 0);
        a.y0 = reinterpret_cast<const double*>(d_y0) + (size_t)b0 * (size_t)p->y0_stride;
        a.y0_stride = p->y0_stride;
        a.data = reinterpret_cast<const char*>(d_data) + (p->data_stride > 0 ? (size_t)b0 * p->data_stride : 0);
        a.t_eval = reinterpret_cast<const double*>(d_te);
        a.rtol_b = p->rtol_b ? reinterpret_cast<const double*>(d_rtol) + b0 : nullptr;
        a.atol_b = p->atol_b ? reinterpret_cast<const double*>(d_atol) + b0 : nullptr;
        a.row_offset = nullptr;
        if (ragged) {
            std::vector<long long>& r = rel[(size_t)c];
            r.resize((size_t)nb + 1);
            for (long long k = 0; k <= nb; k++) r[(size_t)k] = p->row_offset[b0 + k] - p->row_offset[b0];
            CU(cuMemcpyHtoDAsync_v2(d_off[i], r.data(), sizeof(long long) * (size_t)(nb + 1), sk.s));
            a.row_offset = reinterpret_cast<const long long*>(d_off[i]);
        }
        a.t_out = reinterpret_cast<double*>(d_t[i]);
        a.y_out = reinterpret_cast<double*>(d_y[i]);
        a.result = reinterpret_cast<int*>(d_res[i]);
        a.t_event = reinterpret_cast<double*>(d_tev[i]);
        a.y_event = reinterpret_cast<double*>(d_yev[i]);
        if ((rc = launch_ivp(H, P, a, sk.s))) return rc;
        CU(cuEventRecord(ev_kernel[i].e, sk.s));
        if (c >= 1 && (rc = copy_back(c - 1))) return rc;
    }
    if ((rc = copy_back(nchunks - 1))) return rc;
    return drain.wait();
}

extern "C" int b200ode_solve_ivp_args(b200ode_rhs_t H, long long B, int neq, const double* t_span,
                                      long long t_span_stride, const double* y0, long long y0_stride,
                                      const void* data, long long data_stride, int np, int nt,
                                      const double* teval, int n_events, double first_step,
                                      double max_step, double rtol, double atol, long long cap,
                                      const long long* row_offset, double* t_out, double* y_out,
                                      int* result, double* t_event, double* y_event, int device)
{
    b200ode_ivp_problem p;
    memset(&p, 0, sizeof p);
    p.size = sizeof p;
    p.B = B, p.neq = neq;
    p.t_span = t_span, p.t_span_stride = t_span_stride;
    p.y0 = y0, p.y0_stride = y0_stride;
    p.data = data, p.data_stride = data_stride, p.np = np;
    p.nt = nt, p.teval = teval, p.n_events = n_events;
    p.first_step = first_step, p.max_step = max_step, p.rtol = rtol, p.atol = atol;
    p.cap = cap, p.row_offset = row_offset, p.t_out = t_out, p.y_out = y_out;
    p.result = result, p.t_event = t_event, p.y_event = y_event;
    return b200ode_solve_ivp(H, &p, device);
}

// ---------------------------------------------------------------- pinned host memory
extern "C" int b200ode_host_alloc(size_t bytes, void** ptr)
{
    if (!ptr) return fail(B200ODE_EINVAL, "ptr is NULL");
    *ptr = nullptr;
    if (!driver())
        return fail(B200ODE_ENODEV, "CUDA driver (libcuda.so.1) not available");
    CUcontext cur = nullptr;
    CU(cuCtxGetCurrent(&cur));
    if (!cur) {
        CUdevice dev;
        CU(cuDeviceGet(&dev, 0));
        CU(cuDevicePrimaryCtxRetain(&cur, dev));
        CU(cuCtxSetCurrent(cur));
    }
    CU(cuMemHostAlloc(ptr, bytes ? bytes : 1, CU_MEMHOSTALLOC_PORTABLE));
    return 0;
}

extern "C" int b200ode_host_free(void* ptr)
{
    if (!ptr) return 0;
    if (!driver())
        return fail(B200ODE_ENODEV, "CUDA driver (libcuda.so.1) not available");
    CU(cuMemFreeHost(ptr));
    return 0;
}

// ---------------------------------------------------------------- device copies of embedded arrays
namespace {
// makes `device`'s primary context current for the scope of a helper call
struct DeviceCtx {
    CtxRestore restore;
    int rc = 0;
    explicit DeviceCtx(int device)
    {
        if (!driver()) {
            rc = fail(B200ODE_ENODEV, "CUDA driver (libcuda.so.1) not available");
            return;
        }
        if (device < 0) {
            rc = fail(B200ODE_EINVAL, "device = %d", device);
            return;
        }
        CUdevice dev;
        CUcontext ctx = nullptr;
        CUresult r = g_drv.cuDeviceGet(&dev, device);
        if (r == CUDA_SUCCESS) r = g_drv.cuDevicePrimaryCtxRetain(&ctx, dev);
        if (r == CUDA_SUCCESS) r = g_drv.cuCtxSetCurrent(ctx);
        if (r != CUDA_SUCCESS) rc = cu_fail(r, "selecting the device");
    }
};
}  // namespace

extern "C" int b200ode_device_alloc(int device, size_t bytes, void** dptr)
{
    if (!dptr) return fail(B200ODE_EINVAL, "dptr is NULL");
    *dptr = nullptr;
    DeviceCtx c(device);
    if (c.rc) return c.rc;
    CUdeviceptr p = 0;
    CU(cuMemAlloc_v2(&p, bytes ? bytes : 1));
    *dptr = reinterpret_cast<void*>(p);
    return 0;
}

extern "C" int b200ode_device_upload(int device, void* dptr, const void* host, size_t bytes)
{
    if (!dptr || (!host && bytes)) return fail(B200ODE_EINVAL, "NULL argument");
    DeviceCtx c(device);
    if (c.rc) return c.rc;
    if (bytes) CU(cuMemcpyHtoD_v2(reinterpret_cast<CUdeviceptr>(dptr), host, bytes));
    return 0;
}

extern "C" int b200ode_device_free(int device, void* dptr)
{
    if (!dptr) return 0;
    DeviceCtx c(device);
    if (c.rc) return c.rc;
    CU(cuMemFree_v2(reinterpret_cast<CUdeviceptr>(dptr)));
    return 0;
}

// ---------------------------------------------------------------- self test
extern "C" int b200ode_selftest_pow(const double* x_dev, double y, double* out_dev, long long n,
                                    void* stream)
{
    if (!driver())
        return fail(B200ODE_ENODEV, "CUDA driver (libcuda.so.1) not available");
    const B200EmbeddedImage* img = find_image("selftest");
    if (!img) return fail(B200ODE_ENOSYS, "selftest image missing");
    CUcontext cur = nullptr;
    CU(cuCtxGetCurrent(&cur));
    if (!cur) {
        CUdevice dev;
        CU(cuDeviceGet(&dev, 0));
        CU(cuDevicePrimaryCtxRetain(&cur, dev));
        CU(cuCtxSetCurrent(cur));
    }
    CUmodule mod;
    CUfunction fn;
    CU(cuModuleLoadData(&mod, img->begin));
    CU(cuModuleGetFunction(&fn, mod, "b200ode_selftest_pow_kernel"));
    void* params[] = {&x_dev, &y, &out_dev, &n};
    CU(cuLaunchKernel(fn, 1184, 1, 1, 256, 1, 1, 0, static_cast<CUstream>(stream), params, nullptr));
    CU(cuStreamSynchronize(static_cast<CUstream>(stream)));
    CU(cuModuleUnload(mod));
    return 0;
}

extern "C" int b200ode_kernel_info(b200ode_rhs_t H, int info[8])
{
    if (!H || !info) return fail(B200ODE_EINVAL, "NULL argument");
    PerDevice* P;
    int rc = device_state(H, -1, &P);
    if (rc) return rc;
    info[0] = P->regs;
    info[1] = P->local_bytes;
    info[2] = P->smem + P->dyn_smem;
    info[3] = P->threads;
    info[4] = P->blocks_per_sm;
    info[5] = P->sms;
    info[6] = H->last_grid;
    info[7] = H->launches.load();
    return 0;
}
