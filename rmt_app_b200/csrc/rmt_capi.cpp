// rmt_capi.cpp — host side of librmtb200.so (C ABI declared in include/rmt_b200.h).
//
// NVRTC compiles the generated model header + the hand-written kernels for
// sm_100a; the CUDA driver API (dlopen'ed, so that the library loads — and the
// code generator back end works — on a machine without a GPU) loads the cubin
// into the device's primary context and launches the kernels on the caller's
// stream.  There is no CPU implementation of any compute entry point.
#include "../../include/rmt_b200.h"

#include <cuda.h>
#include <nvrtc.h>
#include <nccl.h>          // types only: libnccl.so.2 is dlopen'ed by rmt_comm_* (a process that never shards needs no NCCL)
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...)
{
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

// ---- driver API through dlopen ----------------------------------------------------
#define RMT_STR2(x) #x
#define RMT_STR(x) RMT_STR2(x)
#define RMT_DRIVER_FUNCS(X) \
    X(cuInit) X(cuDeviceGetCount) X(cuDeviceGet) X(cuDevicePrimaryCtxRetain) X(cuCtxSetCurrent) \
    X(cuDeviceGetAttribute) X(cuModuleLoadDataEx) X(cuModuleUnload) X(cuModuleGetFunction) \
    X(cuLaunchKernel) X(cuMemAlloc) X(cuMemFree) X(cuMemcpyHtoDAsync) \
    X(cuMemcpyDtoHAsync) X(cuMemcpyDtoH) X(cuMemsetD8Async) X(cuStreamSynchronize) X(cuFuncSetAttribute) \
    X(cuOccupancyMaxActiveBlocksPerMultiprocessor) X(cuGetErrorString) X(cuEventCreate) X(cuEventRecord) \
    X(cuEventSynchronize) X(cuEventElapsedTime) X(cuEventDestroy) X(cuMemAllocHost) X(cuMemFreeHost) \
    X(cuCtxSynchronize) X(cuStreamCreate) X(cuStreamDestroy_v2)

struct Driver {
    void* lib = nullptr;
#define X(name) decltype(&name) p_##name = nullptr;
    RMT_DRIVER_FUNCS(X)
#undef X
} drv;

std::mutex g_mu;
bool g_driver_ok = false;
int g_device = -1;
CUcontext g_ctx = nullptr;
int g_sm_count = 0;

int load_driver()
{
    if (g_driver_ok) return 0;
    drv.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!drv.lib)
        return fail("rmt_init: cannot load the CUDA driver (libcuda.so.1): %s — this library has no CPU fallback",
                    dlerror());
#define X(name)                                                                                  \
    drv.p_##name = (decltype(&name))dlsym(drv.lib, RMT_STR(name));                               \
    if (!drv.p_##name) return fail("rmt_init: CUDA driver lacks symbol %s", RMT_STR(name));
    RMT_DRIVER_FUNCS(X)
#undef X
    g_driver_ok = true;
    return 0;
}

int cu_fail(CUresult r, const char* what)
{
    const char* s = nullptr;
    if (drv.p_cuGetErrorString) drv.p_cuGetErrorString(r, &s);
    return fail("%s failed: CUDA error %d (%s)", what, (int)r, s ? s : "?");
}
#define CU(call)                                                  \
    do {                                                          \
        CUresult r_ = drv.p_##call;                               \
        if (r_ != CUDA_SUCCESS) return cu_fail(r_, #call);        \
    } while (0)

int ensure_ctx()
{
    if (!g_driver_ok || !g_ctx) return fail("rmt_init(device) has not been called (no CUDA context)");
    CU(cuCtxSetCurrent(g_ctx));
    return 0;
}

// ---- blobs and modules ------------------------------------------------------------
struct Blob {
    std::vector<char> cubin;
    std::string ptx, log;
};

struct Scratch {           // per-launch device scratch: queue counter + z_eval + obj_ref
    CUdeviceptr p = 0;
    size_t bytes = 0;
    CUevent done = nullptr;    // recorded after the launch that uses the slot; waited for before the slot is re-used
    bool in_flight = false;
};

struct Module {
    CUmodule mod = nullptr;
    rmt_module_info info{};
    CUfunction f_setup = nullptr, f_n1_rhs = nullptr, f_n1_jac = nullptr, f_n1_sys = nullptr, f_n1_solve = nullptr;
    CUfunction f_n2_rhs = nullptr, f_n2_solve = nullptr, f_reduce = nullptr, f_peak = nullptr, f_probe = nullptr;
    int solve_blocks_per_sm = 0;
    size_t solve_smem = 0;
    // stage-pipelined N2 module (info.lanes == 0), read from its rmt_n2_meta kernel: reactors per block, work doubles per
    // block and node, work doubles per block
    int wf_reactors = 0;
    long long wf_work_per_node = 0, wf_work_fixed = 0;
    std::mutex mu;             // guards the scratch ring
    std::mutex host_mu;        // serialises the synchronous host-buffer entry points (they share ws / pipe_ws)
    Scratch ring[8];
    int ring_next = 0;
    // grow-only workspaces of the *_host entry points and the reductions
    CUdeviceptr ws[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t ws_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // copy/compute pipeline of rmt_n1_solve_host: two streams, each with its own device buffers
    CUstream pipe_stream[2] = {nullptr, nullptr};
    CUdeviceptr pipe_ws[2][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}};
    size_t pipe_bytes[2][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}};
};

std::map<uint64_t, Blob*> g_blobs;
std::map<uint64_t, Module*> g_modules;
uint64_t g_next_id = 1;

Blob* get_blob(rmt_blob_t b)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_blobs.find(b);
    return it == g_blobs.end() ? nullptr : it->second;
}
Module* get_module(rmt_module_t m)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_modules.find(m);
    return it == g_modules.end() ? nullptr : it->second;
}

int ws_reserve(Module* M, int slot, size_t bytes, CUdeviceptr* out)
{
    if (M->ws_bytes[slot] < bytes) {
        if (M->ws[slot]) { CU(cuCtxSynchronize()); CU(cuMemFree(M->ws[slot])); M->ws[slot] = 0; M->ws_bytes[slot] = 0; }
        size_t want = bytes + bytes/8 + 256;
        CU(cuMemAlloc(&M->ws[slot], want));
        M->ws_bytes[slot] = want;
    }
    *out = M->ws[slot];
    return 0;
}

int pipe_reserve(Module* M, int k, int slot, size_t bytes, CUdeviceptr* out)
{
    if (M->pipe_bytes[k][slot] < bytes) {
        if (M->pipe_ws[k][slot]) { CU(cuCtxSynchronize()); CU(cuMemFree(M->pipe_ws[k][slot])); M->pipe_ws[k][slot] = 0; M->pipe_bytes[k][slot] = 0; }
        size_t want = bytes + bytes/8 + 256;
        CU(cuMemAlloc(&M->pipe_ws[k][slot], want));
        M->pipe_bytes[k][slot] = want;
    }
    *out = M->pipe_ws[k][slot];
    return 0;
}

// Per-launch scratch (work-queue counter, z_eval, objective reference): a ring of 8 slots per module.  Every slot
// carries an event recorded behind the launch that uses it; taking a slot waits for that event, so any number of
// launches may be in flight on any streams (the 9th simply waits for the 1st to finish).  Ring bookkeeping is guarded
// by the module's mutex: concurrent host threads may launch on the same module.
int scratch_get(Module* M, size_t bytes, CUdeviceptr* out, Scratch** slot)
{
    std::lock_guard<std::mutex> lk(M->mu);
    Scratch& s = M->ring[M->ring_next];
    M->ring_next = (M->ring_next + 1) % 8;
    if (s.in_flight) { CU(cuEventSynchronize(s.done)); s.in_flight = false; }
    if (!s.done) CU(cuEventCreate(&s.done, CU_EVENT_DISABLE_TIMING));
    if (s.bytes < bytes) {
        if (s.p) { CU(cuMemFree(s.p)); s.p = 0; s.bytes = 0; }
        size_t want = std::max<size_t>(bytes, 64*1024);
        CU(cuMemAlloc(&s.p, want));
        s.bytes = want;
    }
    *out = s.p;
    *slot = &s;
    return 0;
}

int scratch_launched(Module* M, Scratch* s, CUstream st)
{
    std::lock_guard<std::mutex> lk(M->mu);
    CU(cuEventRecord(s->done, st));
    s->in_flight = true;
    return 0;
}

// host-buffer entry point: ensembles at least this large run as a three-chunk copy/compute pipeline
const int64_t RMT_PIPELINE_MIN_B = 1 << 18;

// default step-size controller: safety, max shrink, max growth, tolerance scale kappa, PI beta,
// initial-step factor (applied to the Hairer-Wanner starting step)
const double RMT_DEFAULT_CTRL[6] = {0.8, 5.0, 6.0, 1.0, 0.08, 0.03};

// diagnostics: step log of one instance (rmt_debug_trace)
double* g_trace_buf = nullptr;
long long g_trace_inst = -1;
int g_trace_cap = 0;

// kernel-parameter mirrors of the device structs in rmt_kernels.cu ---------------------
struct SolveArgsN1 {
    CUdeviceptr consts;
    long long B;
    CUdeviceptr z_eval;
    int n_eval;
    int out_mode;
    double rtol, atol;
    int max_steps;
    int dense;
    CUdeviceptr out, status, stats, queue, obj_ref, obj;
    double ctrl[6];
    CUdeviceptr trace;
    long long trace_inst;
    int trace_cap;
    CUdeviceptr red;
    long long red_offset;
};
static_assert(sizeof(SolveArgsN1) == 192, "SolveArgs layout");

struct SolveArgsN2 {
    CUdeviceptr consts;
    long long B;
    int zNo, tNo;
    double period;
    double rtol, atol;
    int max_steps;
    int out_mode;
    CUdeviceptr out, status, stats, work, queue;
    double ctrl[6];
};
static_assert(sizeof(SolveArgsN2) == 144, "SolveArgsN2 layout");

// RmtInputs { const double* rows; i64 B; int map[NIN]; double u[NIN]; }
std::vector<char> pack_inputs(const Module* M, CUdeviceptr rows, long long B, const int32_t* map, const double* u)
{
    const int nin = M->info.nin;
    size_t off_u = 16 + ((4*(size_t)nin + 7)/8)*8;
    std::vector<char> buf(off_u + 8*(size_t)nin, 0);
    memcpy(&buf[0], &rows, 8);
    memcpy(&buf[8], &B, 8);
    memcpy(&buf[16], map, 4*(size_t)nin);
    memcpy(&buf[off_u], u, 8*(size_t)nin);
    return buf;
}

int check_rows(const Module* M, int32_t n_rows, const int32_t* row_map, const double* uniform, const void* rows)
{
    if (!row_map || !uniform) return fail("rmt_setup: row_map and uniform must not be NULL");
    for (int q = 0; q < M->info.nin; ++q) {
        if (row_map[q] >= n_rows) return fail("rmt_setup: row_map[%d]=%d but only %d rows were passed", q, row_map[q], n_rows);
        if (row_map[q] >= 0 && !rows) return fail("rmt_setup: row_map[%d]=%d but rows is NULL", q, row_map[q]);
    }
    return 0;
}

int launch(CUfunction f, unsigned grid, unsigned block, size_t smem, CUstream st, void** params, const char* name)
{
    if (!f) return fail("%s: kernel not present in this module (compiled for the other model?)", name);
    CUresult r = drv.p_cuLaunchKernel(f, grid, 1, 1, block, 1, 1, (unsigned)smem, st, params, nullptr);
    if (r != CUDA_SUCCESS) return cu_fail(r, name);
    return 0;
}

// ---- NCCL through dlopen (cross-GPU reduce / gather of sharded ensembles) ------------------
#define RMT_NCCL_FUNCS(X) \
    X(ncclGetVersion) X(ncclGetUniqueId) X(ncclCommInitRank) X(ncclCommDestroy) X(ncclAllGather) X(ncclAllReduce) \
    X(ncclGetErrorString)

struct Nccl {
    void* lib = nullptr;
#define X(name) decltype(&name) p_##name = nullptr;
    RMT_NCCL_FUNCS(X)
#undef X
} nccl;
bool g_nccl_ok = false;

int load_nccl()
{
    if (g_nccl_ok) return 0;
    // the soname: when the process already carries an NCCL (e.g. the one PyTorch bundles) this resolves to it
    nccl.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!nccl.lib) return fail("rmt_comm: cannot load libnccl.so.2: %s", dlerror());
#define X(name)                                                                                  \
    nccl.p_##name = (decltype(&name))dlsym(nccl.lib, RMT_STR(name));                             \
    if (!nccl.p_##name) return fail("rmt_comm: libnccl.so.2 lacks symbol %s", RMT_STR(name));
    RMT_NCCL_FUNCS(X)
#undef X
    g_nccl_ok = true;
    return 0;
}

int nccl_fail(ncclResult_t r, const char* what)
{
    return fail("%s failed: NCCL error %d (%s)", what, (int)r, nccl.p_ncclGetErrorString ? nccl.p_ncclGetErrorString(r) : "?");
}
#define NC(call)                                                  \
    do {                                                          \
        ncclResult_t r_ = nccl.p_##call;                          \
        if (r_ != ncclSuccess) return nccl_fail(r_, #call);       \
    } while (0)

struct Comm {
    ncclComm_t comm = nullptr;
    int nranks = 0, rank = 0;
};
std::map<uint64_t, Comm*> g_comms;

Comm* get_comm(rmt_comm_t c)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_comms.find(c);
    return it == g_comms.end() ? nullptr : it->second;
}

}  // namespace

extern "C" {

const char* rmt_last_error(void) { return g_err.c_str(); }
const char* rmt_version(void) { return "rmt_app_b200 0.1 (sm_100a, NVRTC)"; }

int rmt_init(int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (load_driver()) return 1;
    CUresult r = drv.p_cuInit(0);
    if (r != CUDA_SUCCESS) return cu_fail(r, "cuInit");
    int n = 0;
    CU(cuDeviceGetCount(&n));
    if (device < 0 || device >= n) return fail("rmt_init: device %d out of range (%d CUDA devices)", device, n);
    if (g_ctx && g_device == device) { CU(cuCtxSetCurrent(g_ctx)); return 0; }
    if (g_ctx && g_device != device)
        return fail("rmt_init: this process is already bound to device %d (one process per GPU)", g_device);
    CUdevice dev;
    CU(cuDeviceGet(&dev, device));
    CU(cuDevicePrimaryCtxRetain(&g_ctx, dev));
    CU(cuCtxSetCurrent(g_ctx));
    CU(cuDeviceGetAttribute(&g_sm_count, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
    g_device = device;
    return 0;
}

int rmt_device_count(int* count)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (load_driver()) return 1;
    CUresult r = drv.p_cuInit(0);
    if (r != CUDA_SUCCESS) return cu_fail(r, "cuInit");
    CU(cuDeviceGetCount(count));
    return 0;
}

int rmt_shutdown(void)
{
    std::vector<uint64_t> ids;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (auto& kv : g_modules) ids.push_back(kv.first);
    }
    for (uint64_t id : ids) rmt_module_free(id);
    return 0;
}

// ---- NVRTC ---------------------------------------------------------------------------
int rmt_nvrtc_compile(const char* model_src, const char* kernels_src, const char* arch, int block,
                      const char* const* extra_opts, int n_extra_opts, rmt_blob_t* blob_out)
{
    if (!model_src || !kernels_src || !blob_out) return fail("rmt_nvrtc_compile: NULL argument");
    nvrtcProgram prog;
    const char* hdr_src[1] = {model_src};
    const char* hdr_name[1] = {"rmt_model.cuh"};
    nvrtcResult r = nvrtcCreateProgram(&prog, kernels_src, "rmt_kernels.cu", 1, hdr_src, hdr_name);
    if (r != NVRTC_SUCCESS) return fail("nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
    std::string a = std::string("--gpu-architecture=") + (arch && *arch ? arch : "sm_100a");
    std::string b = "-DRMT_BLOCK=" + std::to_string(block > 0 ? block : 128);
    std::vector<const char*> opts = {a.c_str(), b.c_str(), "--std=c++17", "-lineinfo", "--fmad=true",
                                     "-default-device"};
    opts.pop_back();   // "-default-device" is not wanted: all device code is annotated explicitly
    for (int i = 0; i < n_extra_opts; ++i) opts.push_back(extra_opts[i]);
    r = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
    Blob* B = new Blob;
    size_t n = 0;
    nvrtcGetProgramLogSize(prog, &n);
    if (n > 1) { B->log.resize(n); nvrtcGetProgramLog(prog, &B->log[0]); }
    if (r != NVRTC_SUCCESS) {
        fail("NVRTC compilation failed (%s):\n%s", nvrtcGetErrorString(r), B->log.c_str());
        delete B;
        nvrtcDestroyProgram(&prog);
        return 1;
    }
    if (nvrtcGetCUBINSize(prog, &n) != NVRTC_SUCCESS || n == 0) {
        fail("NVRTC produced no cubin for %s (need a real architecture, e.g. sm_100a)", a.c_str());
        delete B;
        nvrtcDestroyProgram(&prog);
        return 1;
    }
    B->cubin.resize(n);
    nvrtcGetCUBIN(prog, B->cubin.data());
    if (nvrtcGetPTXSize(prog, &n) == NVRTC_SUCCESS && n > 1) { B->ptx.resize(n); nvrtcGetPTX(prog, &B->ptx[0]); }
    nvrtcDestroyProgram(&prog);
    std::lock_guard<std::mutex> lk(g_mu);
    uint64_t id = g_next_id++;
    g_blobs[id] = B;
    *blob_out = id;
    return 0;
}

int rmt_blob_data(rmt_blob_t blob, const void** data, size_t* size)
{
    Blob* B = get_blob(blob);
    if (!B) return fail("invalid blob handle");
    *data = B->cubin.data();
    *size = B->cubin.size();
    return 0;
}

const char* rmt_blob_log(rmt_blob_t blob)
{
    Blob* B = get_blob(blob);
    return B ? B->log.c_str() : "";
}

int rmt_blob_ptx(rmt_blob_t blob, const char** ptx, size_t* size)
{
    Blob* B = get_blob(blob);
    if (!B) return fail("invalid blob handle");
    *ptx = B->ptx.c_str();
    *size = B->ptx.size();
    return 0;
}

int rmt_blob_free(rmt_blob_t blob)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_blobs.find(blob);
    if (it == g_blobs.end()) return fail("invalid blob handle");
    delete it->second;
    g_blobs.erase(it);
    return 0;
}

// ---- modules -------------------------------------------------------------------------
int rmt_module_load(const void* cubin, size_t size, rmt_module_t* module_out)
{
    if (ensure_ctx()) return 1;
    if (!cubin || !size) return fail("rmt_module_load: empty image");
    Module* M = new Module;
    CUresult r = drv.p_cuModuleLoadDataEx(&M->mod, cubin, 0, nullptr, nullptr);
    if (r != CUDA_SUCCESS) { delete M; return cu_fail(r, "cuModuleLoadDataEx"); }
    int32_t mv[16] = {0};
    {
        CUfunction fm = nullptr;
        CUdeviceptr dmeta = 0;
        if (drv.p_cuModuleGetFunction(&fm, M->mod, "rmt_meta") != CUDA_SUCCESS || !fm) {
            drv.p_cuModuleUnload(M->mod); delete M; return fail("module has no rmt_meta kernel (not an rmt_app_b200 image)");
        }
        r = drv.p_cuMemAlloc(&dmeta, sizeof mv);
        if (r != CUDA_SUCCESS) { drv.p_cuModuleUnload(M->mod); delete M; return cu_fail(r, "cuMemAlloc"); }
        void* params[] = {&dmeta};
        r = drv.p_cuLaunchKernel(fm, 1, 1, 1, 32, 1, 1, 0, nullptr, params, nullptr);
        if (r == CUDA_SUCCESS) r = drv.p_cuMemcpyDtoH(mv, dmeta, sizeof mv);
        drv.p_cuMemFree(dmeta);
        if (r != CUDA_SUCCESS) { drv.p_cuModuleUnload(M->mod); delete M; return cu_fail(r, "rmt_meta"); }
    }
    rmt_module_info& I = M->info;
    I.model = mv[0]; I.n = mv[1]; I.nc = mv[2]; I.nr = mv[3]; I.nin = mv[4]; I.nconst = mv[5]; I.nkp = mv[6];
    I.stages = mv[7]; I.block = mv[8]; I.iso = mv[9];
    I.flops_rhs_alg = mv[10]; I.flops_rhs_wt = mv[11]; I.flops_jac_alg = mv[12]; I.flops_jac_wt = mv[13];
    I.m = mv[14] > 0 ? mv[14] : mv[1];
    // dynamic models: threads per reactor; 0 = the stage-pipelined mapping (32 reactors x (stages + 1) role warps per block)
    I.lanes = (mv[0] == 2 || mv[0] == 9) ? std::max(mv[15], 0) : 1;
    auto get = [&](const char* name) { CUfunction f = nullptr; drv.p_cuModuleGetFunction(&f, M->mod, name); return f; };
    M->f_setup = get("rmt_setup");
    M->f_n1_rhs = get("rmt_n1_rhs");
    M->f_n1_jac = get("rmt_n1_jac");
    M->f_n1_sys = get("rmt_n1_sys");
    M->f_n1_solve = get("rmt_n1_solve");
    M->f_n2_rhs = get("rmt_n2_rhs");
    M->f_n2_solve = get("rmt_n2_solve");
    M->f_reduce = get("rmt_reduce_partials");
    M->f_peak = get("rmt_dfma_peak");
    M->f_probe = get("rmt_math_probe");
    if (!M->f_setup) { drv.p_cuModuleUnload(M->mod); delete M; return fail("module lacks rmt_setup"); }
    const bool dynamic = I.model == 2 || I.model == 9;       // N2 and its dimensional twin M9: rmt_n2_* entry points
    CUfunction fs = !dynamic ? M->f_n1_solve : M->f_n2_solve;
    if (fs && !dynamic) {
        M->solve_smem = (size_t)I.block*8*((size_t)I.m*I.m + (size_t)I.stages*I.m);
        CU(cuFuncSetAttribute(fs, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)M->solve_smem));
        CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&M->solve_blocks_per_sm, fs, I.block, M->solve_smem));
        if (M->solve_blocks_per_sm < 1) { drv.p_cuModuleUnload(M->mod); delete M; return fail("integrator kernel does not fit on an SM (smem %zu B)", M->solve_smem); }
    } else if (fs) {
        // N2: the substitution's shared-memory record, ((n + 1) n + 3 n + 1) rows of (block + 1) doubles (rmt_n2_solve); M9: none
        // and, with several lanes per reactor, the sweeps' hand-over records [block/lanes][stages + 1][2 n + 4]
        M->solve_smem = I.model == 2 ? 8*((size_t)(I.n + 1)*I.n + 3*(size_t)I.n + 1)*((size_t)I.block + 1) : 0;
        if (I.lanes > 1) M->solve_smem += 8*(size_t)(I.block/I.lanes)*(I.stages + 1)*(2*(size_t)I.n + 4)
                                          + 8*((size_t)(I.block/I.lanes)*8 + 2);      // + the substitution's hand-over vectors
        if (I.lanes == 0) {
            // stage pipeline: the kernel's own layout constants (WF_SMEM_BYTES, reactors per block, work rows)
            CUfunction fm2 = get("rmt_n2_meta");
            if (!fm2) { drv.p_cuModuleUnload(M->mod); delete M; return fail("stage-pipelined module lacks rmt_n2_meta"); }
            int32_t wv[8] = {0};
            CUdeviceptr dm = 0;
            r = drv.p_cuMemAlloc(&dm, sizeof wv);
            if (r == CUDA_SUCCESS) {
                void* params[] = {&dm};
                r = drv.p_cuLaunchKernel(fm2, 1, 1, 1, 32, 1, 1, 0, nullptr, params, nullptr);
                if (r == CUDA_SUCCESS) r = drv.p_cuMemcpyDtoH(wv, dm, sizeof wv);
                drv.p_cuMemFree(dm);
            }
            if (r != CUDA_SUCCESS || wv[1] <= 0) { drv.p_cuModuleUnload(M->mod); delete M; return cu_fail(r, "rmt_n2_meta"); }
            M->solve_smem = (size_t)wv[0]; M->wf_reactors = wv[1]; M->wf_work_per_node = wv[2]; M->wf_work_fixed = wv[3];
        }
        if (M->solve_smem) CU(cuFuncSetAttribute(fs, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)M->solve_smem));
        CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&M->solve_blocks_per_sm, fs, I.block, M->solve_smem));
        if (M->solve_blocks_per_sm < 1) { drv.p_cuModuleUnload(M->mod); delete M; return fail("dynamic-model integrator does not fit on an SM (smem %zu B)", M->solve_smem); }
    }
    std::lock_guard<std::mutex> lk(g_mu);
    uint64_t id = g_next_id++;
    g_modules[id] = M;
    *module_out = id;
    return 0;
}

int rmt_module_get_info(rmt_module_t m, rmt_module_info* info)
{
    Module* M = get_module(m);
    if (!M) return fail("invalid module handle");
    *info = M->info;
    return 0;
}

int rmt_module_free(rmt_module_t m)
{
    Module* M;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_modules.find(m);
        if (it == g_modules.end()) return fail("invalid module handle");
        M = it->second;
        g_modules.erase(it);
    }
    if (g_ctx && drv.p_cuCtxSetCurrent) {
        drv.p_cuCtxSetCurrent(g_ctx);
        drv.p_cuCtxSynchronize();
        for (auto& s : M->ring) { if (s.p) drv.p_cuMemFree(s.p); if (s.done) drv.p_cuEventDestroy(s.done); }
        for (auto& w : M->ws) if (w) drv.p_cuMemFree(w);
        for (int k = 0; k < 2; ++k) {
            for (auto& w : M->pipe_ws[k]) if (w) drv.p_cuMemFree(w);
            if (M->pipe_stream[k]) drv.p_cuStreamDestroy_v2(M->pipe_stream[k]);
        }
        if (M->mod) drv.p_cuModuleUnload(M->mod);
    }
    delete M;
    return 0;
}

// ---- compute entry points ------------------------------------------------------------
int rmt_setup(rmt_module_t m, int64_t B, const double* d_rows, int32_t n_rows, const int32_t* row_map,
              const double* uniform, double* d_consts, void* stream)
{
    Module* M = get_module(m);
    if (!M) return fail("invalid module handle");
    if (ensure_ctx()) return 1;
    if (B <= 0) return fail("rmt_setup: B must be positive");
    if (check_rows(M, n_rows, row_map, uniform, d_rows)) return 1;
    std::vector<char> in = pack_inputs(M, (CUdeviceptr)d_rows, B, row_map, uniform);
    CUdeviceptr c = (CUdeviceptr)d_consts;
    void* params[] = {in.data(), &c};
    return launch(M->f_setup, (unsigned)((B + 127)/128), 128, 0, (CUstream)stream, params, "rmt_setup");
}

int rmt_n1_rhs(rmt_module_t m, int64_t B, const double* d_consts, const double* d_y, double* d_f, void* stream)
{
    Module* M = get_module(m);
    if (!M) return fail("invalid module handle");
    if (ensure_ctx()) return 1;
    CUdeviceptr c = (CUdeviceptr)d_consts, y = (CUdeviceptr)d_y, f = (CUdeviceptr)d_f;
    long long b = B;
    void* params[] = {&c, &b, &y, &f};
    return launch(M->f_n1_rhs, (unsigned)((B + 127)/128), 128, 0, (CUstream)stream, params, "rmt_n1_rhs");
}

int rmt_n1_jac(rmt_module_t m, int64_t B, const double* d_consts, const double* d_y, double* d_f, double* d_J,
               void* stream)
{
    Module* M = get_module(m);
    if (!M) return fail("invalid module handle");
    if (ensure_ctx()) return 1;
    CUdeviceptr c = (CUdeviceptr)d_consts, y = (CUdeviceptr)d_y, f = (CUdeviceptr)d_f, J = (CUdeviceptr)d_J;
    long long b = B;
    void* params[] = {&c, &b, &y, &f, &J};
    return launch(M->f_n1_jac, (unsigned)((B + 127)/128), 128, 0, (CUstream)stream, params, "rmt_n1_jac");
}

int rmt_n1_sys(rmt_module_t m, int64_t B, const double* d_consts, const double* d_y, double* d_g, double* d_A,
               void* stream)
{
    Module* M = get_module(m);
    if (!M) return fail("invalid module handle");
    if (!M->f_n1_sys) return fail("module has no steady-state integrator");
    if (ensure_ctx()) return 1;
    CUdeviceptr c = (CUdeviceptr)d_consts, y = (CUdeviceptr)d_y, g = (CUdeviceptr)d_g, A = (CUdeviceptr)d_A;
    long long b = B;
    void* params[] = {&c, &b, &y, &g, &A};
    return launch(M->f_n1_sys, (unsigned)((B + 127)/128), 128, 0, (CUstream)stream, params, "rmt_n1_sys");
}

static int n1_solve_impl(rmt_module_t m, int64_t B, const double* d_consts, int32_t n_eval, const double* z_eval,
                         double rtol, double atol, int32_t max_steps, int32_t dense, int32_t out_mode,
                         double* d_out, int32_t* d_status, int32_t* d_stats,
                         const double* obj_ref, double* d_obj, double* d_red, int64_t red_offset,
                         const double* ctrl, void* stream);

int rmt_n1_solve(rmt_module_t m, int64_t B, const double* d_consts, int32_t n_eval, const double* z_eval,
                 double rtol, double atol, int32_t max_steps, int32_t dense, int32_t out_mode,
                 double* d_out, int32_t* d_status, int32_t* d_stats,
                 const double* obj_ref, double* d_obj, const double* ctrl, void* stream)
{
    return n1_solve_impl(m, B, d_consts, n_eval, z_eval, rtol, atol, max_steps, dense, out_mode, d_out, d_status, d_stats,
                         obj_ref, d_obj, nullptr, 0, ctrl, stream);
}

int rmt_n1_solve_population(rmt_module_t m, int64_t B, const double* d_consts, double z_end, double rtol, double atol,
                            int32_t max_steps, double* d_out, int32_t* d_status, int32_t* d_stats,
                            const double* obj_ref, double* d_obj, double* d_red, int64_t index_offset,
                            const double* ctrl, void* stream)
{
    if (!obj_ref || !d_obj || !d_red) return fail("rmt_n1_solve_population: obj_ref, d_obj and d_red are required");
    return n1_solve_impl(m, B, d_consts, 1, &z_end, rtol, atol, max_steps, 0, 1, d_out, d_status, d_stats,
                         obj_ref, d_obj, d_red, index_offset, ctrl, stream);
}

static int n1_solve_impl(rmt_module_t m, int64_t B, const double* d_consts, int32_t n_eval, const double* z_eval,
                         double rtol, double atol, int32_t max_steps, int32_t dense, int32_t out_mode,
                         double* d_out, int32_t* d_status, int32_t* d_stats,
                         const double* obj_ref, double* d_obj, double* d_red, int64_t red_offset,
                         const double* ctrl, void* stream)
{
    Module* M = get_module(m);
    if (!M) return fail("invalid module handle");
    if (ensure_ctx()) return 1;
    if (M->info.model == 2 || M->info.model == 9) return fail("rmt_n1_solve: module was generated for a dynamic model (N2/M9)");
    if (B <= 0 || n_eval < 1 || !z_eval) return fail("rmt_n1_solve: need B > 0 and at least one output position");
    for (int e = 1; e < n_eval; ++e)
        if (!(z_eval[e] > z_eval[e - 1])) return fail("rmt_n1_solve: z_eval must be strictly increasing");
    if (!(z_eval[n_eval - 1] > 0.0)) return fail("rmt_n1_solve: the last output position must be > 0");
    if (!(rtol > 0.0) || !(atol >= 0.0)) return fail("rmt_n1_solve: rtol must be > 0 and atol >= 0");
    if ((obj_ref == nullptr) != (d_obj == nullptr)) return fail("rmt_n1_solve: obj_ref and d_obj go together");
    if (!d_consts || !d_out || !d_status) return fail("rmt_n1_solve: d_consts, d_out and d_status are required (d_stats may be NULL)");
    CUstream st = (CUstream)stream;
    const int n = M->info.n;
    size_t need = 64 + 8*(size_t)n_eval + 8*(size_t)n;
    CUdeviceptr scr;
    Scratch* slot = nullptr;
    if (scratch_get(M, need, &scr, &slot)) return 1;
    CU(cuMemsetD8Async(scr, 0, 64, st));
    CU(cuMemcpyHtoDAsync(scr + 64, z_eval, 8*(size_t)n_eval, st));
    if (obj_ref) CU(cuMemcpyHtoDAsync(scr + 64 + 8*(size_t)n_eval, obj_ref, 8*(size_t)n, st));
    SolveArgsN1 a;
    a.consts = (CUdeviceptr)d_consts; a.B = B; a.z_eval = scr + 64; a.n_eval = n_eval; a.out_mode = out_mode;
    a.rtol = rtol; a.atol = atol; a.max_steps = max_steps > 0 ? max_steps : 100000; a.dense = dense ? 1 : 0;
    a.out = (CUdeviceptr)d_out; a.status = (CUdeviceptr)d_status; a.stats = (CUdeviceptr)d_stats; a.queue = scr;
    a.obj_ref = obj_ref ? scr + 64 + 8*(size_t)n_eval : 0; a.obj = (CUdeviceptr)d_obj;
    a.trace = (CUdeviceptr)g_trace_buf; a.trace_inst = g_trace_inst; a.trace_cap = g_trace_cap;
    a.red = (CUdeviceptr)d_red; a.red_offset = red_offset;
    for (int k = 0; k < 6; ++k) a.ctrl[k] = ctrl ? ctrl[k] : RMT_DEFAULT_CTRL[k];
    if (!(a.ctrl[0] > 0.0 && a.ctrl[0] <= 1.0) || !(a.ctrl[1] > 1.0) || !(a.ctrl[2] > 1.0) || !(a.ctrl[3] > 0.0))
        return fail("rmt_n1_solve: controller needs 0 < safety <= 1, max shrink > 1, max growth > 1, kappa > 0");
    if (!(a.ctrl[5] > 0.0)) a.ctrl[5] = RMT_DEFAULT_CTRL[5];
    const int block = M->info.block;
    // One block per SM; lanes pull reactors from the queue, a warp at a time.  An ensemble smaller than one resident
    // wave is spread over all SMs (a warp's worth of reactors per block and up) rather than packed into a few of
    // them: the integrator is bound by the FP64 pipe of the SM it runs on (8 192 reactors: 22 full blocks use 15 %
    // of the GPU's FP64 pipes, 148 blocks of two busy warps all of them).
    long long want = (B + 31)/32;
    long long cap = (long long)g_sm_count*M->solve_blocks_per_sm;
    unsigned grid = (unsigned)std::max<long long>(1, std::min(want, cap));
    void* params[] = {&a};
    if (launch(M->f_n1_solve, grid, block, M->solve_smem, st, params, "rmt_n1_solve")) return 1;
    return scratch_launched(M, slot, st);
}

int rmt_n1_solve_host(rmt_module_t m, int64_t B, const double* h_rows, int32_t n_rows, const int32_t* row_map,
                      const double* uniform, int32_t n_eval, const double* z_eval,
                      double rtol, double atol, int32_t max_steps, int32_t dense, int32_t out_mode,
                      double* h_out, int32_t* h_status, int32_t* h_stats,
                      const double* obj_ref, double* h_obj, const double* ctrl)
{
    Module* M = get_module(m);
    if (!M) return fail("invalid module handle");
    if (ensure_ctx()) return 1;
    if (B <= 0) return fail("rmt_n1_solve_host: B must be positive");
    std::lock_guard<std::mutex> host_lk(M->host_mu);
    if (check_rows(M, n_rows, row_map, uniform, h_rows)) return 1;
    if (!h_out || !h_status) return fail("rmt_n1_solve_host: h_out and h_status are required");
    const int n = M->info.n;
    const int rows = out_mode == 2 ? 2*n + M->info.nc : n;
    if (B >= RMT_PIPELINE_MIN_B && n_rows > 0) {
        // Large host-side ensembles: three chunks (1/8, 3/4, 1/8 of the reactors, 1024-aligned) on two streams, so
        // that the copies of one chunk run under the integrator kernel of another and the blocks of the next kernel
        // fill the tail of the previous one.  Every reactor is an independent solve: same bits as one launch.
        for (int k = 0; k < 2; ++k)
            if (!M->pipe_stream[k]) CU(cuStreamCreate(&M->pipe_stream[k], CU_STREAM_NON_BLOCKING));
        const int64_t c1 = std::max<int64_t>(1024, ((B/8 + 512)/1024)*1024);
        const int64_t cuts[4] = {0, c1, B - c1, B};
        for (int c = 0; c < 3; ++c) {
            const int64_t b0 = cuts[c], Bc = cuts[c + 1] - cuts[c];
            const int k = c % 2;
            CUstream st = M->pipe_stream[k];
            CUdeviceptr d_rows, d_consts, d_out, d_status, d_stats, d_obj = 0;
            if (pipe_reserve(M, k, 0, 8*(size_t)n_rows*Bc, &d_rows)) return 1;
            if (pipe_reserve(M, k, 1, 8*(size_t)M->info.nconst*Bc, &d_consts)) return 1;
            if (pipe_reserve(M, k, 2, 8*(size_t)n_eval*rows*Bc, &d_out)) return 1;
            if (pipe_reserve(M, k, 3, 4*(size_t)Bc, &d_status)) return 1;
            if (pipe_reserve(M, k, 4, 16*(size_t)Bc, &d_stats)) return 1;
            if (obj_ref && pipe_reserve(M, k, 5, 8*(size_t)Bc, &d_obj)) return 1;
            for (int r = 0; r < n_rows; ++r)
                CU(cuMemcpyHtoDAsync(d_rows + 8*(size_t)r*Bc, h_rows + (size_t)r*B + b0, 8*(size_t)Bc, st));
            if (rmt_setup(m, Bc, (const double*)d_rows, n_rows, row_map, uniform, (double*)d_consts, st)) return 1;
            if (rmt_n1_solve(m, Bc, (const double*)d_consts, n_eval, z_eval, rtol, atol, max_steps, dense, out_mode,
                             (double*)d_out, (int32_t*)d_status, (int32_t*)d_stats, obj_ref, (double*)d_obj, ctrl, st)) return 1;
            for (int64_t q = 0; q < (int64_t)n_eval*rows; ++q)
                CU(cuMemcpyDtoHAsync(h_out + (size_t)q*B + b0, d_out + 8*(size_t)q*Bc, 8*(size_t)Bc, st));
            CU(cuMemcpyDtoHAsync(h_status + b0, d_status, 4*(size_t)Bc, st));
            if (h_stats)
                for (int q = 0; q < 4; ++q)
                    CU(cuMemcpyDtoHAsync(h_stats + (size_t)q*B + b0, d_stats + 4*(size_t)q*Bc, 4*(size_t)Bc, st));
            if (obj_ref && h_obj) CU(cuMemcpyDtoHAsync(h_obj + b0, d_obj, 8*(size_t)Bc, st));
        }
        CU(cuStreamSynchronize(M->pipe_stream[0]));
        CU(cuStreamSynchronize(M->pipe_stream[1]));
        return 0;
    }
    const size_t bytes_rows = 8*(size_t)std::max(n_rows, 0)*B, bytes_consts = 8*(size_t)M->info.nconst*B;
    const size_t bytes_out = 8*(size_t)n_eval*rows*B;
    CUdeviceptr d_rows = 0, d_consts, d_out, d_status, d_stats, d_obj = 0;
    if (n_rows > 0 && ws_reserve(M, 0, bytes_rows, &d_rows)) return 1;
    if (ws_reserve(M, 1, bytes_consts, &d_consts)) return 1;
    if (ws_reserve(M, 2, bytes_out, &d_out)) return 1;
    if (ws_reserve(M, 3, 4*(size_t)B, &d_status)) return 1;
    if (ws_reserve(M, 4, 16*(size_t)B, &d_stats)) return 1;
    if (obj_ref && ws_reserve(M, 5, 8*(size_t)B, &d_obj)) return 1;
    CUstream st = nullptr;
    if (n_rows > 0) CU(cuMemcpyHtoDAsync(d_rows, h_rows, bytes_rows, st));
    if (rmt_setup(m, B, (const double*)d_rows, n_rows, row_map, uniform, (double*)d_consts, st)) return 1;
    if (rmt_n1_solve(m, B, (const double*)d_consts, n_eval, z_eval, rtol, atol, max_steps, dense, out_mode,
                     (double*)d_out, (int32_t*)d_status, (int32_t*)d_stats, obj_ref, (double*)d_obj, ctrl, st)) return 1;
    CU(cuMemcpyDtoHAsync(h_out, d_out, bytes_out, st));
    CU(cuMemcpyDtoHAsync(h_status, d_status, 4*(size_t)B, st));
    if (h_stats) CU(cuMemcpyDtoHAsync(h_stats, d_stats, 16*(size_t)B, st));
    if (obj_ref && h_obj) CU(cuMemcpyDtoHAsync(h_obj, d_obj, 8*(size_t)B, st));
    CU(cuStreamSynchronize(st));
    return 0;
}

int rmt_n2_rhs(rmt_module_t m, int64_t B, int32_t zNo, const double* d_consts, const double* d_y, double* d_f,
               void* stream)
{
    Module* M = get_module(m);
    if (!M) return fail("invalid module handle");
    if (ensure_ctx()) return 1;
    if (M->info.model != 2 && M->info.model != 9) return fail("rmt_n2_rhs: module was generated for a steady-state model");
    if (zNo < 2) return fail("rmt_n2_rhs: zNo must be >= 2");
    CUdeviceptr c = (CUdeviceptr)d_consts, y = (CUdeviceptr)d_y, f = (CUdeviceptr)d_f;
    long long b = B;
    int z = zNo;
    void* params[] = {&c, &b, &z, &y, &f};
    return launch(M->f_n2_rhs, (unsigned)((B + 63)/64), 64, 0, (CUstream)stream, params, "rmt_n2_rhs");
}

// threads of the N2 integrator launch: `lanes` threads per reactor, whole blocks, at most one resident wave
static int64_t n2_slots(const Module* M, int64_t B)
{
    const int block = M->info.block;
    if (M->info.lanes == 0) {             // stage pipeline: a block serves wf_reactors reactors at a time
        long long wantb = (B + M->wf_reactors - 1)/M->wf_reactors;
        long long capb = (long long)g_sm_count*std::max(M->solve_blocks_per_sm, 1);
        return (int64_t)std::max<long long>(1, std::min(wantb, capb))*block;
    }
    long long want = (B*M->info.lanes + block - 1)/block;
    long long cap = (long long)g_sm_count*std::max(M->solve_blocks_per_sm, 1);
    return (int64_t)std::max<long long>(1, std::min(want, cap))*block;
}

int64_t rmt_n2_work_doubles(rmt_module_t m, int64_t B, int32_t zNo)
{
    Module* M = get_module(m);
    if (!M) { fail("invalid module handle"); return -1; }
    if (B <= 0 || zNo < 2) { fail("rmt_n2_work_doubles: need B > 0 and zNo >= 2"); return -1; }
    const int64_t n = M->info.n, s = M->info.stages;
    // per node and integrator thread: y_n, y_{n+1}, s stage vectors, W_kk^{-1} (n x n), upwind / pressure
    // coupling vectors (3n), d E/d P
    // (M9 adds the velocity march: two more coupling vectors, one more gradient vector, four scalars)
    if (M->info.lanes == 0)               // stage pipeline, per block: y_n / y_{n+1} and the ring of stage vectors (rmt_n2_meta)
        return (n2_slots(M, B)/M->info.block)*(M->wf_work_per_node*zNo + M->wf_work_fixed);
    const int64_t rows = n*(5 + s + n) + 1 + (M->info.model == 9 ? 3*n + 4 : 0);
    const int64_t groups = (zNo + M->info.lanes - 1)/M->info.lanes;      // node groups, one node per lane
    // the two state rows exist per node group; the rest is scratch of the group being processed
    return (groups*2*n + (rows - 2*n))*n2_slots(M, B);
}

int rmt_n2_solve(rmt_module_t m, int64_t B, int32_t zNo, int32_t tNo, double period, const double* d_consts,
                 double rtol, double atol, int32_t max_steps, int32_t out_mode,
                 double* d_out, int32_t* d_status, int32_t* d_stats, double* d_work, const double* ctrl, void* stream)
{
    Module* M = get_module(m);
    if (!M) return fail("invalid module handle");
    if (ensure_ctx()) return 1;
    if (M->info.model != 2 && M->info.model != 9) return fail("rmt_n2_solve: module was generated for a steady-state model");
    if (B <= 0 || zNo < 2 || tNo < 1 || !(period > 0.0)) return fail("rmt_n2_solve: need B > 0, zNo >= 2, tNo >= 1, period > 0");
    if (!(rtol > 0.0) || !(atol >= 0.0)) return fail("rmt_n2_solve: rtol must be > 0 and atol >= 0");
    if (!d_work) return fail("rmt_n2_solve: d_work is required (rmt_n2_work_doubles)");
    if (!d_consts || !d_out || !d_status) return fail("rmt_n2_solve: d_consts, d_out and d_status are required (d_stats may be NULL)");
    if (M->info.n > 15) return fail("rmt_n2_solve: more than 15 unknowns per node are not supported (pivot packing)");
    CUstream st = (CUstream)stream;
    CUdeviceptr scr;
    Scratch* slot = nullptr;
    if (scratch_get(M, 64, &scr, &slot)) return 1;
    CU(cuMemsetD8Async(scr, 0, 64, st));
    SolveArgsN2 a;
    a.consts = (CUdeviceptr)d_consts; a.B = B; a.zNo = zNo; a.tNo = tNo; a.period = period;
    a.rtol = rtol; a.atol = atol; a.max_steps = max_steps > 0 ? max_steps : 1000000; a.out_mode = out_mode;
    a.out = (CUdeviceptr)d_out; a.status = (CUdeviceptr)d_status; a.stats = (CUdeviceptr)d_stats;
    a.work = (CUdeviceptr)d_work; a.queue = scr;
    for (int k = 0; k < 6; ++k) a.ctrl[k] = ctrl ? ctrl[k] : RMT_DEFAULT_CTRL[k];
    if (!(a.ctrl[0] > 0.0 && a.ctrl[0] <= 1.0) || !(a.ctrl[1] > 1.0) || !(a.ctrl[2] > 1.0) || !(a.ctrl[3] > 0.0))
        return fail("rmt_n2_solve: controller needs 0 < safety <= 1, max shrink > 1, max growth > 1, kappa > 0");
    if (!(a.ctrl[5] > 0.0)) a.ctrl[5] = RMT_DEFAULT_CTRL[5];
    const int block = M->info.block;
    unsigned grid = (unsigned)(n2_slots(M, B)/block);
    void* params[] = {&a};
    if (launch(M->f_n2_solve, grid, block, M->solve_smem, st, params, "rmt_n2_solve")) return 1;
    return scratch_launched(M, slot, st);
}

int rmt_reduce_objective(rmt_module_t m, int64_t n, const double* d_obj, int64_t index_offset,
                         double* h_sum, double* h_min, int64_t* h_argmin, void* stream)
{
    Module* M = get_module(m);
    if (!M) return fail("invalid module handle");
    if (ensure_ctx()) return 1;
    if (n <= 0) return fail("rmt_reduce_objective: n must be positive");
    std::lock_guard<std::mutex> host_lk(M->host_mu);
    CUstream st = (CUstream)stream;
    const unsigned blocks = (unsigned)std::min<long long>(256, (n + 255)/256);
    CUdeviceptr part;
    if (ws_reserve(M, 6, 3*8*(size_t)(blocks + 1), &part)) return 1;
    CUdeviceptr v = (CUdeviceptr)d_obj, ps = part, pm = part + 8*(size_t)(blocks + 1), pa = part + 16*(size_t)(blocks + 1);
    long long nn = n, off = index_offset;
    void* params[] = {&v, &nn, &off, &ps, &pm, &pa};
    if (launch(M->f_reduce, blocks, 256, 0, st, params, "rmt_reduce_partials")) return 1;
    std::vector<double> hs(blocks), hm(blocks);
    std::vector<long long> ha(blocks);
    CU(cuMemcpyDtoHAsync(hs.data(), ps, 8*(size_t)blocks, st));
    CU(cuMemcpyDtoHAsync(hm.data(), pm, 8*(size_t)blocks, st));
    CU(cuMemcpyDtoHAsync(ha.data(), pa, 8*(size_t)blocks, st));
    CU(cuStreamSynchronize(st));
    double s = 0.0, mn = hm[0];
    long long am = ha[0];
    for (unsigned b = 0; b < blocks; ++b) {           // fixed order: deterministic
        s += hs[b];
        if (hm[b] < mn || (hm[b] == mn && ha[b] >= 0 && (am < 0 || ha[b] < am))) { mn = hm[b]; am = ha[b]; }
    }
    if (h_sum) *h_sum = s;
    if (h_min) *h_min = mn;
    if (h_argmin) *h_argmin = am;
    return 0;
}

// ---- sharded ensembles: cross-GPU gather / reduce (NCCL over NVLink) -----------------------
int rmt_comm_unique_id(void* id_out, size_t id_bytes)
{
    if (!id_out || id_bytes < sizeof(ncclUniqueId)) return fail("rmt_comm_unique_id: need a buffer of RMT_COMM_ID_BYTES (%zu) bytes", sizeof(ncclUniqueId));
    if (load_nccl()) return 1;
    ncclUniqueId id;
    NC(ncclGetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return 0;
}

int rmt_comm_init(int32_t nranks, int32_t rank, const void* id, size_t id_bytes, rmt_comm_t* comm_out)
{
    if (!id || id_bytes < sizeof(ncclUniqueId) || !comm_out) return fail("rmt_comm_init: NULL / short id or NULL comm_out");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail("rmt_comm_init: rank %d of %d", rank, nranks);
    if (ensure_ctx()) return 1;                       // rmt_init(device) first: one process per GPU
    if (load_nccl()) return 1;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    Comm* C = new Comm;
    C->nranks = nranks; C->rank = rank;
    ncclResult_t r = nccl.p_ncclCommInitRank(&C->comm, nranks, uid, rank);
    if (r != ncclSuccess) { delete C; return nccl_fail(r, "ncclCommInitRank"); }
    std::lock_guard<std::mutex> lk(g_mu);
    uint64_t h = g_next_id++;
    g_comms[h] = C;
    *comm_out = h;
    return 0;
}

int rmt_comm_info(rmt_comm_t comm, int32_t* nranks, int32_t* rank, int32_t* nccl_version)
{
    Comm* C = get_comm(comm);
    if (!C) return fail("invalid communicator handle");
    if (nranks) *nranks = C->nranks;
    if (rank) *rank = C->rank;
    if (nccl_version) { int v = 0; NC(ncclGetVersion(&v)); *nccl_version = v; }
    return 0;
}

int rmt_comm_allgather(rmt_comm_t comm, const double* d_send, double* d_recv, int64_t count, void* stream)
{
    Comm* C = get_comm(comm);
    if (!C) return fail("invalid communicator handle");
    if (count <= 0 || !d_send || !d_recv) return fail("rmt_comm_allgather: need count > 0 and device buffers");
    if (ensure_ctx()) return 1;
    NC(ncclAllGather(d_send, d_recv, (size_t)count, ncclFloat64, C->comm, (cudaStream_t)stream));
    return 0;
}

int rmt_comm_allreduce(rmt_comm_t comm, const double* d_send, double* d_recv, int64_t count, int32_t op, void* stream)
{
    Comm* C = get_comm(comm);
    if (!C) return fail("invalid communicator handle");
    if (count <= 0 || !d_send || !d_recv) return fail("rmt_comm_allreduce: need count > 0 and device buffers");
    if (op < 0 || op > 2) return fail("rmt_comm_allreduce: op 0 = sum, 1 = min, 2 = max");
    if (ensure_ctx()) return 1;
    const ncclRedOp_t ops[3] = {ncclSum, ncclMin, ncclMax};
    NC(ncclAllReduce(d_send, d_recv, (size_t)count, ncclFloat64, ops[op], C->comm, (cudaStream_t)stream));
    return 0;
}

int rmt_comm_free(rmt_comm_t comm)
{
    Comm* C;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_comms.find(comm);
        if (it == g_comms.end()) return fail("invalid communicator handle");
        C = it->second;
        g_comms.erase(it);
    }
    if (C->comm && nccl.p_ncclCommDestroy) nccl.p_ncclCommDestroy(C->comm);
    delete C;
    return 0;
}

int rmt_math_probe(rmt_module_t m, int32_t n, const double* d_x, double* d_out, void* stream)
{
    Module* M = get_module(m);
    if (!M) return fail("invalid module handle");
    if (ensure_ctx()) return 1;
    CUdeviceptr x = (CUdeviceptr)d_x, o = (CUdeviceptr)d_out;
    int nn = n;
    void* params[] = {&x, &nn, &o};
    return launch(M->f_probe, (unsigned)((n + 127)/128), 128, 0, (CUstream)stream, params, "rmt_math_probe");
}

int rmt_debug_trace(double* d_trace, int32_t cap, int64_t instance)
{
    g_trace_buf = d_trace; g_trace_cap = d_trace ? cap : 0; g_trace_inst = d_trace ? instance : -1;
    return 0;
}

int rmt_fp64_peak(rmt_module_t m, int32_t iters, int32_t repeats, double* tflops_out)
{
    Module* M = get_module(m);
    if (!M) return fail("invalid module handle");
    if (ensure_ctx()) return 1;
    if (!M->f_peak) return fail("module lacks rmt_dfma_peak");
    std::lock_guard<std::mutex> host_lk(M->host_mu);
    const unsigned blocks = (unsigned)g_sm_count*8, threads = 256;
    CUdeviceptr out;
    if (ws_reserve(M, 7, 8*(size_t)blocks*threads, &out)) return 1;
    int it = iters > 0 ? iters : 4096;
    double seed = 1.0;
    void* params[] = {&out, &it, &seed};
    CUevent e0, e1;
    CU(cuEventCreate(&e0, 0));
    CU(cuEventCreate(&e1, 0));
    double best = 0.0;
    for (int r = 0; r < std::max(repeats, 1) + 1; ++r) {
        CU(cuEventRecord(e0, nullptr));
        if (launch(M->f_peak, blocks, threads, 0, nullptr, params, "rmt_dfma_peak")) return 1;
        CU(cuEventRecord(e1, nullptr));
        CU(cuEventSynchronize(e1));
        float ms = 0.f;
        CU(cuEventElapsedTime(&ms, e0, e1));
        double fl = 2.0*8.0*(double)it*(double)blocks*threads;
        if (r > 0) best = std::max(best, fl/(ms*1e-3)/1e12);
    }
    CU(cuEventDestroy(e0));
    CU(cuEventDestroy(e1));
    *tflops_out = best;
    return 0;
}

}  // extern "C"
