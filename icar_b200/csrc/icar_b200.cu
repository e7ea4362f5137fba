// C ABI of libicar_b200 (include/icar_b200.h): the host side that mirrors the reference's
// advect / mp / halo_send / halo_retrieve call surface and drives the sm_100a kernels.
//
// Reference host logic restated here:
//   src/physics/advection_driver.f90:51-77   advect dispatch
//   src/physics/adv_mpdata.f90:463-524       mpdata(): winds, scalar list
//   src/physics/mp_driver.f90:398-772        mp(): update interval, tiles, precipitation
//   src/objects/domain_obj.f90:104-155       halo_send / halo_retrieve
//   src/objects/grid_obj.f90:116-255         decomposition bounds
//   src/main/time_step.f90:512-529,547       the physics part of step()
#include "../../include/icar_b200.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "halo.cuh"
#include "linear_winds.cuh"
#include "mp_simple.cuh"
#include "mp_thompson.cuh"
#include "mpdata.cuh"
#include "step_loop.cuh"

using namespace icar;

namespace {

std::string g_create_error;

enum Shape { CELL, USTAG, VSTAG, D2 };

Shape var_shape(int var) {
    switch (var) {
        case ICAR_VAR_U:
        case ICAR_VAR_JACOBIAN_U: return USTAG;
        case ICAR_VAR_V:
        case ICAR_VAR_JACOBIAN_V: return VSTAG;
        case ICAR_VAR_PRECIPITATION:
        case ICAR_VAR_SNOWFALL:
        case ICAR_VAR_GRAUPEL: return D2;
        default: return CELL;
    }
}

// dynamically bound NCCL (so the library loads on a host without NCCL / without a GPU).  NCCL is plumbing here:
// the all-gather of the arena handles at icar_comm_init and the one-double co_min of update_dt.  The halo exchange
// itself is written by kernels into peer-mapped memory (halo.cu).
struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define BIND(field, sym)                                                     \
    *(void**)(&field) = dlsym(h, sym);                                       \
    if (!field) { err = std::string("libnccl lacks ") + sym; return false; }
        BIND(GetUniqueId, "ncclGetUniqueId")
        BIND(CommInitRank, "ncclCommInitRank")
        BIND(CommDestroy, "ncclCommDestroy")
        BIND(AllGather, "ncclAllGather")
        BIND(AllReduce, "ncclAllReduce")
        BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
        return true;
    }
};
Nccl g_nccl;

}  // namespace

struct icar_ctx {
    int device = 0;
    icar_grid grid{};
    icar_options opt{};
    Geom g{};
    int ld2 = 0;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev_packed = nullptr, ev_comm = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    void* vars[ICAR_VAR_COUNT] = {};
    bool uploaded[ICAR_VAR_COUNT] = {};
    // module arrays of adv_mpdata (U_m, V_m, W_m) + per-call temporaries
    float *U = nullptr, *V = nullptr, *W = nullptr, *Wn = nullptr, *G = nullptr;
    float *qa = nullptr, *qb = nullptr, *u2 = nullptr, *v2 = nullptr, *w2 = nullptr;
    // fused path (mpdata_fused.cu): coefficient fields, per-scalar pass-1 scratch and output buffers
    bool fused = false;
    FusedCoef coef{};
    float* fq2[ICAR_MAX_SCALARS] = {};
    float* fout[ICAR_MAX_SCALARS] = {};
    // scalars advect() moves (adv_mpdata.f90:512-522 order) and halo_send exchanges (domain_obj.f90:106-119
    // order), as requested by the microphysics scheme's *_var_request (mp_simple.f90:116-118, mp_driver.f90:128-131)
    int nadv = 0, adv[ICAR_MAX_SCALARS] = {};
    int nexch = 0, exch[ICAR_MAX_EXCHANGED] = {};
    // Thompson's module array SR (mp_driver.f90:440)
    float* sr2d = nullptr;
    cudaStream_t aux = nullptr;      // the ring strips of mp(halo) run here, next to mp(subset)
    cudaEvent_t ev_fork = nullptr, ev_aux = nullptr, ev_winds = nullptr;
    bool aux_pending = false;        // ring strips still running on the aux stream
    bool winds_pending = false;      // the wind-only part of advect() still running on the aux stream (next to mp())
    unsigned long long* diag = nullptr;
    std::vector<void*> scratch;      // leaf-call staging buffers (float or double cells)
    std::vector<size_t> scratch_bytes;
    double model_time = 0.0, last_model_time = -999.0;
    long long launches = 0;
    std::string err;
    // multi-image
    ncclComm_t comm = nullptr;
    int image = 1, nimages = 1;
    // The advected scalars (two buffers each: advect() writes the other one) and the exchange flags live in ONE
    // allocation, the arena, so that a neighbouring image can map it (CUDA IPC, or the plain pointer inside one
    // process) and store its boundary cells straight into this image's ghost cells (halo.cu).
    void* arena = nullptr;
    size_t slot_bytes = 0;                   // one buffer of one scalar, rounded up to 256 bytes
    float* home[ICAR_MAX_SCALARS][2] = {};   // the two buffers of advected scalar n (adv[] order)
    struct Peer {
        void* base = nullptr;                // the neighbour's arena as this process sees it
        bool ipc = false;                    // opened with cudaIpcOpenMemHandle (to be closed)
        int nx = 0, ny = 0, ld = 0;
        size_t slot_bytes = 0;
    } peer[2];                               // [0] lower image (west / south), [1] upper image
    bool linked = false;                     // icar_comm_init / icar_comm_init_local has connected the neighbours
    bool exchange_in_flight = false;
    bool winds_f64_valid = false;
    // per-kernel profile (icar_profile_enable / icar_profile_read): event pairs queued on the launching stream
    bool prof_on = false;
    struct ProfRec { int kid; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[16] = {};
    long long prof_n[16] = {};
    // module state of mp_thompson (constants + the 29 lookup tables), filled by icar_thompson_init
    bool th_ready = false;
    ThConst thc{};
    ThTables tht{};
    double* th_tab[TH_NTABLES] = {};
    double *th_bins = nullptr, *th_aux = nullptr;
    unsigned long long* th_diag = nullptr;     // [0] columns that took the no_micro exit
    float* dqdt[ICAR_VAR_COUNT] = {};          // var%dqdt_3d of the forced variables (icar_upload_tendency)
    // compute_dt: 4 maxima (uint images of floats) | 1 double for the co_min | nz dz_levels
    unsigned* cfl_max = nullptr;
    double* cfl_dt = nullptr;
    float* cfl_dz = nullptr;
    // icar_physics_steps: two consecutive steps captured as one CUDA graph (two, because advect() swaps each
    // field with its output buffer, so the pointer state has period 2)
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        float dt = 0.f;
        bool runs[2] = {false, false};       // does mp() run in step A / B (update_interval, mp_driver.f90:703)
        float mp_dt[2] = {0.f, 0.f};         // and with which mp_dt (:704)
        void* vars[ICAR_VAR_COUNT] = {};     // the buffer each variable lived in when the pair was captured
        long long launches = 0;              // kernels of this library inside the pair
        bool unsupported = false;            // pointer state did not return after two steps: stay eager
    } sgraph;
    long long eager_steps = 0;               // completed icar_physics_step calls with dt > 1e-3 (lazy allocations done)
    LinWinds* lw = nullptr;                  // module state of linear_theory_winds (icar_linear_winds_init)
};

namespace {

int fail(icar_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

// kernel families of the per-kernel profile
enum ProfKid { K_MP_SIMPLE = 0, K_MP_THOMPSON, K_WIND_NORMALISE, K_FACE_COEF, K_UPWIND_PASS1, K_CORRECTOR,
               K_STAGED_ADVECT3D, K_HALO_PUSH, K_HALO_WAIT, K_COUNT };
const char* const kProfNames[K_COUNT] = {"k_mp_simple", "k_mp_thompson", "k_wind_normalise",
                                         "k_face_coefficients", "k_upwind_all", "k_corrector_fused",
                                         "staged_advect3d", "k_halo_push", "k_halo_wait"};

cudaEvent_t prof_event(icar_ctx* c) {
    if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
// brackets the launches queued on `st` between construction and destruction
struct ProfScope {
    icar_ctx* c; int kid; cudaStream_t st; cudaEvent_t a = nullptr;
    ProfScope(icar_ctx* c_, int kid_, cudaStream_t st_) : c(c_), kid(kid_), st(st_) {
        if (c->prof_on) { a = prof_event(c); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (a) { cudaEvent_t b = prof_event(c); cudaEventRecord(b, st); c->prof_pending.push_back({kid, a, b}); }
    }
};

#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return fail(ctx, ICAR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

#define NC(call)                                                                         \
    do {                                                                                 \
        ncclResult_t r__ = (call);                                                       \
        if (r__ != ncclSuccess)                                                          \
            return fail(ctx, ICAR_ERR_COMM, std::string(#call) + ": " + g_nccl.GetErrorString(r__)); \
    } while (0)

size_t elem_size(int var) { return var_shape(var) == D2 ? sizeof(double) : sizeof(float); }

// rows (x-lines) and row geometry of a variable
void var_geometry(const icar_ctx* c, int var, int* width, int* pitch, size_t* rows) {
    const Geom& g = c->g;
    switch (var_shape(var)) {
        case CELL:  *width = g.nx;     *pitch = g.ld;   *rows = (size_t)g.nz * g.ny; break;
        case USTAG: *width = g.nx + 1; *pitch = g.ldu;  *rows = (size_t)g.nz * g.ny; break;
        case VSTAG: *width = g.nx;     *pitch = g.ld;   *rows = (size_t)g.nz * (g.ny + 1); break;
        case D2:    *width = g.nx;     *pitch = c->ld2; *rows = (size_t)g.ny; break;
    }
}

constexpr size_t kArenaHeader = 256;      // HF_COUNT exchange flags in front of the scalar buffers
static_assert(kArenaHeader >= HF_COUNT * sizeof(unsigned), "flag block");

size_t cell_elems(const icar_ctx* c) { return (size_t)c->g.ld * c->g.nz * (c->g.ny + 1); }

int alloc_cells(icar_ctx* ctx, float** p) {
    CU(cudaMalloc((void**)p, cell_elems(ctx) * sizeof(float)));
    CU(cudaMemsetAsync(*p, 0, cell_elems(ctx) * sizeof(float), ctx->stream));
    return ICAR_OK;
}

// staging buffer #slot with room for one padded field of `esize`-byte elements
int get_scratch(icar_ctx* ctx, size_t slot, size_t esize, void** out) {
    size_t need = (size_t)std::max(ctx->g.ldu, ctx->g.ld) * ctx->g.nz * (ctx->g.ny + 1) * esize;
    if (ctx->scratch.size() <= slot) { ctx->scratch.resize(slot + 1, nullptr); ctx->scratch_bytes.resize(slot + 1, 0); }
    if (ctx->scratch_bytes[slot] < need) {
        if (ctx->scratch[slot]) CU(cudaFree(ctx->scratch[slot]));
        ctx->scratch[slot] = nullptr;
        CU(cudaMalloc(&ctx->scratch[slot], need));
        CU(cudaMemsetAsync(ctx->scratch[slot], 0, need, ctx->stream));
        ctx->scratch_bytes[slot] = need;
    }
    *out = ctx->scratch[slot];
    return ICAR_OK;
}

// host (dense, reference layout) <-> device (padded rows)
// (a dense field -- pitch == width, e.g. nx = 512 -- goes as ONE linear copy: a 2-D copy of 2 KB rows
// runs at pageable-memory speed even from pinned memory)
int copy_in(icar_ctx* ctx, void* dev, const void* host, int width, int pitch, size_t rows, size_t es) {
    if (pitch == width)
        CU(cudaMemcpyAsync(dev, host, (size_t)width * es * rows, cudaMemcpyHostToDevice, ctx->stream));
    else
        CU(cudaMemcpy2DAsync(dev, (size_t)pitch * es, host, (size_t)width * es, (size_t)width * es, rows,
                             cudaMemcpyHostToDevice, ctx->stream));
    return ICAR_OK;
}
int copy_out(icar_ctx* ctx, void* host, const void* dev, int width, int pitch, size_t rows, size_t es) {
    if (pitch == width)
        CU(cudaMemcpyAsync(host, dev, (size_t)width * es * rows, cudaMemcpyDeviceToHost, ctx->stream));
    else
        CU(cudaMemcpy2DAsync(host, (size_t)width * es, dev, (size_t)pitch * es, (size_t)width * es, rows,
                             cudaMemcpyDeviceToHost, ctx->stream));
    return ICAR_OK;
}

// advect(): the order of adv_mpdata.f90:512-522 restricted to options%vars_to_advect
const int kAdvectedSimple[5] = {ICAR_VAR_WATER_VAPOR, ICAR_VAR_CLOUD_WATER, ICAR_VAR_RAIN_IN_AIR,
                                ICAR_VAR_SNOW_IN_AIR, ICAR_VAR_POTENTIAL_TEMPERATURE};
const int kAdvectedThompson[9] = {ICAR_VAR_WATER_VAPOR, ICAR_VAR_CLOUD_WATER, ICAR_VAR_RAIN_IN_AIR,
                                  ICAR_VAR_SNOW_IN_AIR, ICAR_VAR_POTENTIAL_TEMPERATURE, ICAR_VAR_CLOUD_ICE,
                                  ICAR_VAR_GRAUPEL_IN_AIR, ICAR_VAR_ICE_NUMBER, ICAR_VAR_RAIN_NUMBER};
// order of domain%halo_send (domain_obj.f90:109-124) restricted to the allocated scalars
// (tests/test_reference_source.py::test_live_domain_halo_calls executes that routine against these lists)
const int kExchangedSimple[5] = {ICAR_VAR_WATER_VAPOR, ICAR_VAR_POTENTIAL_TEMPERATURE,
                                 ICAR_VAR_CLOUD_WATER, ICAR_VAR_RAIN_IN_AIR, ICAR_VAR_SNOW_IN_AIR};
const int kExchangedThompson[9] = {ICAR_VAR_WATER_VAPOR, ICAR_VAR_POTENTIAL_TEMPERATURE, ICAR_VAR_CLOUD_WATER,
                                   ICAR_VAR_CLOUD_ICE, ICAR_VAR_ICE_NUMBER, ICAR_VAR_RAIN_IN_AIR,
                                   ICAR_VAR_RAIN_NUMBER, ICAR_VAR_SNOW_IN_AIR, ICAR_VAR_GRAUPEL_IN_AIR};

bool check_dims(const icar_ctx* c, int ims, int ime, int kms, int kme, int jms, int jme) {
    return (ime - ims + 1) == c->g.nx && (kme - kms + 1) == c->g.nz && (jme - jms + 1) == c->g.ny;
}

int need_uploaded(icar_ctx* ctx, std::initializer_list<int> vars, const char* who) {
    for (int v : vars)
        if (!ctx->uploaded[v])
            return fail(ctx, ICAR_ERR_INVALID, std::string(who) + ": domain variable " + std::to_string(v) +
                                                   " was never uploaded (icar_upload)");
    return ICAR_OK;
}

// process_subdomain's scheme call (mp_driver.f90:398-607) on up to four 1-based inclusive tiles
// {its, ite, jts, jte}, one launch
int run_mp_tiles(icar_ctx* ctx, float mp_dt, const int (*t)[4], int ntiles, int kts, int kte, cudaStream_t st) {
    double* acc_rain = (double*)ctx->vars[ICAR_VAR_PRECIPITATION];
    double* acc_snow = (double*)ctx->vars[ICAR_VAR_SNOWFALL];
    const icar_grid& gr = ctx->grid;
    MpTiles tiles;
    tiles.n = 0;
    for (int a = 0; a < ntiles; ++a) {
        if (t[a][1] < t[a][0] || t[a][3] < t[a][2]) continue;
        tiles.its[tiles.n] = t[a][0] - gr.ims; tiles.ite[tiles.n] = t[a][1] - gr.ims;
        tiles.jts[tiles.n] = t[a][2] - gr.jms; tiles.jte[tiles.n] = t[a][3] - gr.jms;
        // strips narrower than a warp in x but long in y: run the lanes along y
        tiles.along_y[tiles.n] = ((t[a][1] - t[a][0] + 1) < 16 && (t[a][3] - t[a][2] + 1) > (t[a][1] - t[a][0] + 1)) ? 1 : 0;
        ++tiles.n;
    }
    if (tiles.n == 0) return ICAR_OK;
    if (ctx->opt.microphysics == 1) {              // process_subdomain's mp_gt_driver call, mp_driver.f90:415-440
        if (!ctx->th_ready) return fail(ctx, ICAR_ERR_INVALID, "icar_mp: thompson_init has not run");
        for (int a = 0; a < tiles.n; ++a) {        // i_end = MIN(ite, ide-1), j_end = MIN(jte, jde-1)  (mp_thompson.f90:821-822)
            tiles.ite[a] = std::min(tiles.ite[a], gr.ide - 1 - gr.ims);
            tiles.jte[a] = std::min(tiles.jte[a], gr.jde - 1 - gr.jms);
            tiles.along_y[a] = ((tiles.ite[a] - tiles.its[a] + 1) < 4 && (tiles.jte[a] - tiles.jts[a]) > (tiles.ite[a] - tiles.its[a])) ? 1 : 0;
        }
        ProfScope ps(ctx, K_MP_THOMPSON, st);
        int nl = th_driver_launch(
            ctx->g, ctx->thc, ctx->tht, (float*)ctx->vars[ICAR_VAR_WATER_VAPOR], (float*)ctx->vars[ICAR_VAR_CLOUD_WATER],
            (float*)ctx->vars[ICAR_VAR_RAIN_IN_AIR], (float*)ctx->vars[ICAR_VAR_CLOUD_ICE],
            (float*)ctx->vars[ICAR_VAR_SNOW_IN_AIR], (float*)ctx->vars[ICAR_VAR_GRAUPEL_IN_AIR],
            (float*)ctx->vars[ICAR_VAR_ICE_NUMBER], (float*)ctx->vars[ICAR_VAR_RAIN_NUMBER],
            (float*)ctx->vars[ICAR_VAR_POTENTIAL_TEMPERATURE], (const float*)ctx->vars[ICAR_VAR_EXNER],
            (const float*)ctx->vars[ICAR_VAR_PRESSURE], (const float*)ctx->vars[ICAR_VAR_DZ_MASS], mp_dt, nullptr,
            /*RAINNCV = this_precip: never written by the reference*/ nullptr, nullptr, nullptr, nullptr, nullptr,
            ctx->sr2d, ctx->ld2, tiles, kts - gr.kms, kte - gr.kms, ctx->th_diag, st, acc_rain, acc_snow,
            (double*)ctx->vars[ICAR_VAR_GRAUPEL]);
        if (nl < 0) return fail(ctx, ICAR_ERR_INVALID, "mp_thompson: more than 64 levels are not built");
        ctx->launches += nl;
        CU(cudaGetLastError());
        return ICAR_OK;
    }
    if (ctx->opt.microphysics == 2) {
        float c2s = std::exp(-1.0f * (1.0f / 2000.0f) * mp_dt);    // mp_simple.f90:619
        float c2r = std::exp(-1.0f * (1.0f / 500.0f) * mp_dt);     // mp_simple.f90:620
        ProfScope ps(ctx, K_MP_SIMPLE, st);
        int nl = mp_simple_launch(
            ctx->g, (const float*)ctx->vars[ICAR_VAR_PRESSURE],
            (float*)ctx->vars[ICAR_VAR_POTENTIAL_TEMPERATURE], (const float*)ctx->vars[ICAR_VAR_EXNER],
            (const float*)ctx->vars[ICAR_VAR_DENSITY], (float*)ctx->vars[ICAR_VAR_WATER_VAPOR],
            (float*)ctx->vars[ICAR_VAR_CLOUD_WATER], (float*)ctx->vars[ICAR_VAR_RAIN_IN_AIR],
            (float*)ctx->vars[ICAR_VAR_SNOW_IN_AIR], nullptr, nullptr, ctx->ld2, mp_dt,
            (const float*)ctx->vars[ICAR_VAR_DZ_MASS], tiles, kts - gr.kms, kte - gr.kms, c2r, c2s, ctx->diag, st,
            acc_rain, acc_snow);
        if (nl < 0) return fail(ctx, ICAR_ERR_INVALID, "mp_simple: more than 80 levels are not built");
        ctx->launches += nl;
        CU(cudaGetLastError());
        return ICAR_OK;
    }
    return fail(ctx, ICAR_ERR_INVALID, "microphysics scheme " + std::to_string(ctx->opt.microphysics) +
                                           " is not built (kMP_THOMPSON = 1 and kMP_SB04 = 2 are)");
}

// Makes the compute stream wait for the strips that mp(halo) / the multi-image step left running on the
// auxiliary stream.  (The per-call precipitation needs no joining: the mp kernels add it to the REAL*8
// accumulators themselves, the strips and the interior tile own disjoint columns.)
int join_aux(icar_ctx* ctx) {
    if (!ctx->aux_pending) return ICAR_OK;
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_aux, 0));
    ctx->aux_pending = false;
    return ICAR_OK;
}

// The wind-only part of advect() (wind normalisation, face coefficients) reads u, v, w and the metric and writes the
// module arrays U_m, V_m, W_m and the coefficient planes: nothing mp() touches.  icar_physics_step queues it on the
// auxiliary stream, where these HBM-bound kernels run next to the issue-bound microphysics kernel, and joins it here
// before the scalars are advected.
int join_winds(icar_ctx* ctx) {
    if (!ctx->winds_pending) return ICAR_OK;
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_winds, 0));
    ctx->winds_pending = false;
    return ICAR_OK;
}

}  // namespace

// ============================================================================ lifecycle

extern "C" void icar_default_options(icar_options* opt) {
    if (!opt) return;
    opt->advection = 2;
    opt->microphysics = 2;
    opt->mpdata_order = 2;
    opt->flux_corrected_transport = 1;
    opt->advect_density = 0;
    opt->update_interval = 0;
    opt->top_mp_level = 0;
    opt->dx = 4000.0f;
    opt->exact_arithmetic = 0;
}

extern "C" int icar_grid_xslab(int nx_global, int nz, int ny_global, int nimages, int image,
                               int halo_size, icar_grid* out) {
    if (!out || nimages < 1 || image < 1 || image > nimages || nx_global < nimages || halo_size < 1)
        return ICAR_ERR_INVALID;
    // my_n / my_start, grid_obj.f90:116-138 with ximages = nimages, yimages = 1
    int base_n = nx_global / nimages, rem = nx_global % nimages;
    int n_local = base_n + (image <= rem ? 1 : 0);
    int start = (image - 1) * base_n + (image - 1 < rem ? image - 1 : rem) + 1;
    bool west = (image == 1), east = (image == nimages);
    icar_grid g;
    g.ims = start; g.ime = start + n_local - 1;
    g.jms = 1; g.jme = ny_global;
    g.kms = 1; g.kme = nz; g.kts = 1; g.kte = nz;
    g.ids = 1; g.ide = nx_global; g.jds = 1; g.jde = ny_global; g.kds = 1; g.kde = nz;
    // update_with_halos, grid_obj.f90:209-255
    g.ims -= west ? 0 : halo_size;
    g.ime += east ? 0 : halo_size;
    g.its = g.ims + (west ? 1 : halo_size);
    g.ite = g.ime - (east ? 1 : halo_size);
    g.jts = g.jms + 1;
    g.jte = g.jme - 1;
    g.halo_size = halo_size;
    *out = g;
    return ICAR_OK;
}

extern "C" int icar_grid_yslab(int nx_global, int nz, int ny_global, int nimages, int image,
                               int halo_size, icar_grid* out) {
    if (!out || nimages < 1 || image < 1 || image > nimages || ny_global < nimages || halo_size < 1)
        return ICAR_ERR_INVALID;
    // my_n / my_start, grid_obj.f90:116-138 with ximages = 1, yimages = nimages
    int base_n = ny_global / nimages, rem = ny_global % nimages;
    int n_local = base_n + (image <= rem ? 1 : 0);
    int start = (image - 1) * base_n + (image - 1 < rem ? image - 1 : rem) + 1;
    bool south = (image == 1), north = (image == nimages);
    icar_grid g;
    g.jms = start; g.jme = start + n_local - 1;
    g.ims = 1; g.ime = nx_global;
    g.kms = 1; g.kme = nz; g.kts = 1; g.kte = nz;
    g.ids = 1; g.ide = nx_global; g.jds = 1; g.jde = ny_global; g.kds = 1; g.kde = nz;
    g.jms -= south ? 0 : halo_size;                  // update_with_halos, grid_obj.f90:209-255
    g.jme += north ? 0 : halo_size;
    g.jts = g.jms + (south ? 1 : halo_size);
    g.jte = g.jme - (north ? 1 : halo_size);
    g.its = g.ims + 1;
    g.ite = g.ime - 1;
    g.halo_size = halo_size;
    *out = g;
    return ICAR_OK;
}

extern "C" int icar_create(int device, const icar_grid* grid, const icar_options* opt, icar_ctx** out) {
    icar_ctx* ctx = nullptr;
    if (!grid || !opt || !out) return fail(nullptr, ICAR_ERR_INVALID, "icar_create: null argument");
    int nx = grid->ime - grid->ims + 1, nz = grid->kme - grid->kms + 1, ny = grid->jme - grid->jms + 1;
    if (nx < 3 || ny < 3 || nz < 1)
        return fail(nullptr, ICAR_ERR_INVALID, "icar_create: tile must be at least 3 x 1 x 3 cells");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, ICAR_ERR_CUDA,
                    std::string("icar_create: no usable CUDA device (") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                        "); this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, ICAR_ERR_INVALID, "icar_create: bad device index");
    CU(cudaSetDevice(device));
    ctx = new icar_ctx();
    ctx->device = device;
    ctx->grid = *grid;
    ctx->opt = *opt;
    ctx->g.nx = nx; ctx->g.nz = nz; ctx->g.ny = ny;
    ctx->g.ld = (nx + 31) / 32 * 32;
    ctx->g.ldu = nx + 1;                       // staggered rows are dense: one linear H2D copy
    ctx->ld2 = ctx->g.ld;
    auto bail = [&](int code) { std::string m = ctx->err; icar_destroy(ctx); g_create_error = m; return code; };
#define CUB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); return bail(ICAR_ERR_CUDA); } } while (0)
    // the exchange and the boundary strips are short and on the critical path of a multi-image step: they get
    // the highest stream priority so that their CTAs are scheduled ahead of the interior tile's
    int prio_lo = 0, prio_hi = 0;
    CUB(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CUB(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, prio_hi));
    CUB(cudaEventCreateWithFlags(&ctx->ev_packed, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
    CUB(cudaEventCreate(&ctx->ev_t0));
    CUB(cudaEventCreate(&ctx->ev_t1));
    if (opt->microphysics == 1) {
        ctx->nadv = 9; std::copy(kAdvectedThompson, kAdvectedThompson + 9, ctx->adv);
        ctx->nexch = 9; std::copy(kExchangedThompson, kExchangedThompson + 9, ctx->exch);
    } else {
        ctx->nadv = 5; std::copy(kAdvectedSimple, kAdvectedSimple + 5, ctx->adv);
        ctx->nexch = 5; std::copy(kExchangedSimple, kExchangedSimple + 5, ctx->exch);
    }
    {   // the arena: exchange flags, then two buffers per advected scalar
        ctx->slot_bytes = (cell_elems(ctx) * sizeof(float) + 255) / 256 * 256;
        const size_t bytes = kArenaHeader + (size_t)2 * ctx->nadv * ctx->slot_bytes;
        CUB(cudaMalloc(&ctx->arena, bytes));
        CUB(cudaMemsetAsync(ctx->arena, 0, bytes, ctx->stream));
        for (int n = 0; n < ctx->nadv; ++n)
            for (int b = 0; b < 2; ++b)
                ctx->home[n][b] = (float*)((char*)ctx->arena + kArenaHeader + (size_t)(2 * n + b) * ctx->slot_bytes);
    }
    for (int v = 0; v < ICAR_VAR_COUNT; ++v) {
        const int* a = std::find(ctx->adv, ctx->adv + ctx->nadv, v);
        if (a != ctx->adv + ctx->nadv) {              // an advected scalar: lives in the arena
            ctx->vars[v] = ctx->home[a - ctx->adv][0];
            ctx->fout[a - ctx->adv] = ctx->home[a - ctx->adv][1];
            continue;
        }
        int width, pitch; size_t rows;
        var_geometry(ctx, v, &width, &pitch, &rows);
        size_t bytes = (var_shape(v) == D2) ? (size_t)pitch * rows * sizeof(double)
                       : (var_shape(v) == USTAG) ? (size_t)ctx->g.ldu * nz * ny * sizeof(float)
                                                 : cell_elems(ctx) * sizeof(float);
        CUB(cudaMalloc(&ctx->vars[v], bytes));
        CUB(cudaMemsetAsync(ctx->vars[v], 0, bytes, ctx->stream));
    }
    float** cellbufs[] = {&ctx->U, &ctx->V, &ctx->W, &ctx->Wn, &ctx->qa, &ctx->qb, &ctx->u2, &ctx->v2, &ctx->w2};
    for (float** p : cellbufs)
        if (alloc_cells(ctx, p) != ICAR_OK) return bail(ICAR_ERR_CUDA);
    if (opt->advect_density && alloc_cells(ctx, &ctx->G) != ICAR_OK) return bail(ICAR_ERR_CUDA);
    ctx->fused = (opt->advection == 2 && opt->mpdata_order == 2 && opt->flux_corrected_transport &&
                  !opt->advect_density && !opt->exact_arithmetic && fused_plan(ctx->g).ok);
    if (ctx->fused) {
        float** cf[] = {&ctx->coef.t0x, &ctx->coef.cbx, &ctx->coef.ccx, &ctx->coef.t0y, &ctx->coef.cby, &ctx->coef.ccy,
                        &ctx->coef.ggy, &ctx->coef.t0z, &ctx->coef.cbz, &ctx->coef.ccz};
        for (float** p : cf)
            if (alloc_cells(ctx, p) != ICAR_OK) return bail(ICAR_ERR_CUDA);
        for (int n = 0; n < ctx->nadv; ++n)
            if (alloc_cells(ctx, &ctx->fq2[n]) != ICAR_OK) return bail(ICAR_ERR_CUDA);
    }
    CUB(cudaMalloc((void**)&ctx->sr2d, (size_t)ctx->ld2 * ny * sizeof(float)));
    CUB(cudaMemsetAsync(ctx->sr2d, 0, (size_t)ctx->ld2 * ny * sizeof(float), ctx->stream));
    CUB(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CUB(cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, prio_hi));
    CUB(cudaEventCreateWithFlags(&ctx->ev_aux, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&ctx->ev_winds, cudaEventDisableTiming));
    CUB(cudaMalloc((void**)&ctx->diag, 3 * sizeof(unsigned long long)));
    CUB(cudaMemsetAsync(ctx->diag, 0, 3 * sizeof(unsigned long long), ctx->stream));
    CUB(cudaStreamSynchronize(ctx->stream));
#undef CUB
    if (opt->microphysics == 1) {          // mp_init -> thompson_init(options%mp_options), mp_driver.f90:80-83
        int rc = icar_thompson_init(ctx, nullptr);
        if (rc) return bail(rc);
    }
    *out = ctx;
    return ICAR_OK;
}

extern "C" int icar_destroy(icar_ctx* ctx) {
    if (!ctx) return ICAR_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    if (ctx->aux) cudaStreamSynchronize(ctx->aux);
    if (ctx->sgraph.exec) cudaGraphExecDestroy(ctx->sgraph.exec);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    for (auto& pr : ctx->peer) if (pr.base && pr.ipc) cudaIpcCloseMemHandle(pr.base);
    for (void* p : ctx->vars)         // (the advected scalars live in the arena)
        if (p && !((char*)p >= (char*)ctx->arena && ctx->arena && (char*)p < (char*)ctx->arena + kArenaHeader + (size_t)2 * ctx->nadv * ctx->slot_bytes)) cudaFree(p);
    if (ctx->arena) cudaFree(ctx->arena);
    float* bufs[] = {ctx->U, ctx->V, ctx->W, ctx->Wn, ctx->G, ctx->qa, ctx->qb, ctx->u2, ctx->v2, ctx->w2, ctx->sr2d};
    for (float* p : bufs) if (p) cudaFree(p);
    if (ctx->diag) cudaFree(ctx->diag);
    for (double* t : ctx->th_tab) if (t) cudaFree(t);
    if (ctx->th_bins) cudaFree(ctx->th_bins);
    if (ctx->th_aux) cudaFree(ctx->th_aux);
    if (ctx->th_diag) cudaFree(ctx->th_diag);
    for (float* t : ctx->dqdt) if (t) cudaFree(t);
    if (ctx->cfl_max) cudaFree(ctx->cfl_max);
    if (ctx->cfl_dt) cudaFree(ctx->cfl_dt);
    if (ctx->cfl_dz) cudaFree(ctx->cfl_dz);
    {
        float* cf[] = {ctx->coef.t0x, ctx->coef.cbx, ctx->coef.ccx, ctx->coef.t0y, ctx->coef.cby, ctx->coef.ccy,
                       ctx->coef.ggy, ctx->coef.t0z, ctx->coef.cbz, ctx->coef.ccz};
        for (float* p : cf) if (p) cudaFree(p);
        for (int n = 0; n < ICAR_MAX_SCALARS; ++n) if (ctx->fq2[n]) cudaFree(ctx->fq2[n]);
    }
    for (void* p : ctx->scratch) if (p) cudaFree(p);
    lw_destroy(ctx->lw);
    for (auto& r : ctx->prof_pending) { if (r.kid >= 0) cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_aux) cudaEventDestroy(ctx->ev_aux);
    if (ctx->ev_winds) cudaEventDestroy(ctx->ev_winds);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    if (ctx->ev_packed) cudaEventDestroy(ctx->ev_packed);
    if (ctx->ev_comm) cudaEventDestroy(ctx->ev_comm);
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    delete ctx;
    return ICAR_OK;
}

extern "C" int icar_device_count(int* n) {
    if (!n) return ICAR_ERR_INVALID;
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { (void)cudaGetLastError(); c = 0; }
    *n = c;
    return ICAR_OK;
}

extern "C" const char* icar_last_error(const icar_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

// ============================================================================ resident state

extern "C" int icar_upload(icar_ctx* ctx, int var, const void* host) {
    if (ctx && ctx->exchange_in_flight)
        return fail(ctx, ICAR_ERR_INVALID, "icar_upload between icar_halo_send and icar_halo_retrieve: the neighbours are writing this image's ghost cells");
    if (!ctx || !host || var < 0 || var >= ICAR_VAR_COUNT) return fail(ctx, ICAR_ERR_INVALID, "icar_upload: bad argument");
    CU(cudaSetDevice(ctx->device));
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    int width, pitch; size_t rows;
    var_geometry(ctx, var, &width, &pitch, &rows);
    int rc = copy_in(ctx, ctx->vars[var], host, width, pitch, rows, elem_size(var));
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->uploaded[var] = true;
    return ICAR_OK;
}

extern "C" int icar_download(icar_ctx* ctx, int var, void* host) {
    if (ctx && ctx->exchange_in_flight)
        return fail(ctx, ICAR_ERR_INVALID, "icar_download between icar_halo_send and icar_halo_retrieve: the neighbours are writing this image's ghost cells");
    if (!ctx || !host || var < 0 || var >= ICAR_VAR_COUNT) return fail(ctx, ICAR_ERR_INVALID, "icar_download: bad argument");
    CU(cudaSetDevice(ctx->device));
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    int width, pitch; size_t rows;
    var_geometry(ctx, var, &width, &pitch, &rows);
    int rc = copy_out(ctx, host, ctx->vars[var], width, pitch, rows, elem_size(var));
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return ICAR_OK;
}

extern "C" int icar_device_ptr(icar_ctx* ctx, int var, void** dptr, long long* pitch_elems) {
    if (!ctx || var < 0 || var >= ICAR_VAR_COUNT || !dptr) return fail(ctx, ICAR_ERR_INVALID, "icar_device_ptr: bad argument");
    int width, pitch; size_t rows;
    var_geometry(ctx, var, &width, &pitch, &rows);
    *dptr = ctx->vars[var];
    if (pitch_elems) *pitch_elems = pitch;
    return ICAR_OK;
}

extern "C" int icar_set_model_time(icar_ctx* ctx, double seconds) {
    if (!ctx) return ICAR_ERR_INVALID;
    ctx->model_time = seconds;
    return ICAR_OK;
}
extern "C" int icar_get_model_time(icar_ctx* ctx, double* seconds) {
    if (!ctx || !seconds) return ICAR_ERR_INVALID;
    *seconds = ctx->model_time;
    return ICAR_OK;
}

extern "C" int icar_synchronize(icar_ctx* ctx) {
    if (!ctx) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->comm_stream));
    return ICAR_OK;
}

extern "C" int icar_mp_diagnostics(icar_ctx* ctx, long long out[3]) {
    if (!ctx || !out) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    unsigned long long h[3];
    CU(cudaMemcpyAsync(h, ctx->diag, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 3; ++i) out[i] = (long long)h[i];
    return ICAR_OK;
}

extern "C" long long icar_kernel_launches(const icar_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int icar_fused_kernel_name(icar_ctx* ctx, char name[64]) {
    if (!ctx || !name) return ICAR_ERR_INVALID;
    name[0] = 0;
    if (!ctx->fused) return ICAR_OK;                       // the staged kernels serve this configuration
    CU(cudaSetDevice(ctx->device));
    const FusedPlan p = fused_plan(ctx->g);
    std::snprintf(name, 64, "k_corrector_fused<%d, %d, %d>", p.cpt / 2, p.nzp, p.stages);
    return ICAR_OK;
}

extern "C" int icar_selftest_division(icar_ctx* ctx, const float* a, const float* b, float* packed, float* ieee, long long n) {
    if (!ctx || !a || !b || !packed || !ieee || n <= 0 || (n & 1)) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    float* d = nullptr;
    const size_t bytes = (size_t)n * sizeof(float);
    CU(cudaMalloc((void**)&d, 4 * bytes));
    cudaError_t e = cudaMemcpyAsync(d, a, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, b, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        ctx->launches += selftest_division(d, d + n, d + 2 * n, d + 3 * n, n, ctx->stream);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(packed, d + 2 * n, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ieee, d + 3 * n, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    CU(e);
    return ICAR_OK;
}

extern "C" int icar_timer_start(icar_ctx* ctx) {
    if (!ctx) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev_t0, ctx->stream));
    return ICAR_OK;
}

extern "C" int icar_timer_stop(icar_ctx* ctx, float* elapsed_ms) {
    if (!ctx || !elapsed_ms) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    CU(cudaEventRecord(ctx->ev_t1, ctx->stream));
    CU(cudaEventSynchronize(ctx->ev_t1));
    CU(cudaEventElapsedTime(elapsed_ms, ctx->ev_t0, ctx->ev_t1));
    return ICAR_OK;
}

// ============================================================================ advect

// part 0: the whole advect(); part 1: what depends on the winds only (wind normalisation, face
// coefficients) -- a multi-image step queues it before it waits for the halo exchange; part 2: the scalars.
int advect_impl(icar_ctx* ctx, float dt, int part, cudaStream_t winds_stream) {
    if (!ctx) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (part != 1) { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    // part 1 may run on another stream (icar_physics_step: next to mp()); parts 0 and 2 run on the compute stream
    cudaStream_t ws = (part == 1 && winds_stream) ? winds_stream : ctx->stream;
    if (ctx->opt.advection != 1 && ctx->opt.advection != 2)
        return ICAR_OK;   // advect() does nothing for other values (advection_driver.f90:67-71)
    int rc = need_uploaded(ctx, {ICAR_VAR_U, ICAR_VAR_V, ICAR_VAR_W, ICAR_VAR_JACOBIAN, ICAR_VAR_JACOBIAN_U,
                                 ICAR_VAR_JACOBIAN_V, ICAR_VAR_JACOBIAN_W, ICAR_VAR_ADVECTION_DZ}, "icar_advect");
    if (rc) return rc;
    if (ctx->opt.advect_density && (rc = need_uploaded(ctx, {ICAR_VAR_DENSITY}, "icar_advect"))) return rc;
    const Geom& g = ctx->g;
    const float* rho = ctx->opt.advect_density ? (const float*)ctx->vars[ICAR_VAR_DENSITY] : nullptr;
    const float* jaco = (const float*)ctx->vars[ICAR_VAR_JACOBIAN];
    const float* dz = (const float*)ctx->vars[ICAR_VAR_ADVECTION_DZ];
    const bool mpdata = (ctx->opt.advection == 2);
    // adv_mpdata.f90:500-506 / advect.f90:344-347
    if (part != 2) {
        ProfScope ps(ctx, K_WIND_NORMALISE, ws);
        ctx->launches += staged_wind_normalise<float>(
            g, (const float*)ctx->vars[ICAR_VAR_U], (const float*)ctx->vars[ICAR_VAR_V],
            (const float*)ctx->vars[ICAR_VAR_W], (const float*)ctx->vars[ICAR_VAR_JACOBIAN_U],
            (const float*)ctx->vars[ICAR_VAR_JACOBIAN_V], (const float*)ctx->vars[ICAR_VAR_JACOBIAN_W], rho, dt,
            ctx->opt.dx, mpdata ? 2 : 1, ctx->U, ctx->V, ctx->W, ws, ctx->fused ? dz : nullptr,
            ctx->fused ? ctx->Wn : nullptr);
    }
    const int order = mpdata ? ctx->opt.mpdata_order : 1;
    const float* G = jaco;          // jaco*rho with rho == 1 is jaco itself
    if (order > 1 && !ctx->fused) {
        if (part != 2) ctx->launches += staged_wn_G<float>(g, ctx->W, dz, jaco, rho, ctx->Wn, rho ? ctx->G : nullptr, ws);
        if (rho) G = ctx->G;
    }
    if (ctx->fused) {
        ScalarSet ss;
        ss.n = ctx->nadv;
        for (int n = 0; n < ctx->nadv; ++n) {
            ss.q[n] = (const float*)ctx->vars[ctx->adv[n]];
            ss.q2[n] = ctx->fq2[n];
            ss.out[n] = ctx->fout[n];
        }
        cudaEvent_t marks[4];
        if (ctx->prof_on) for (cudaEvent_t& m : marks) m = prof_event(ctx);
        // a slab stores only the cells it owns: its ghost cells are refilled by the next exchange before anything reads them
        const icar_grid& gr = ctx->grid;
        const int own[4] = {gr.its - gr.ims, gr.ite - gr.ims + 1, gr.jts - gr.jms, gr.jte - gr.jms + 1};
        const int nl = fused_advect_all(g, ss, ctx->U, ctx->V, ctx->W, ctx->Wn, jaco, dz, ctx->coef, ws,
                                        ctx->prof_on ? marks : nullptr, part, ctx->nimages > 1 ? own : nullptr);
        if (ctx->prof_on) {      // consecutive marks delimit the three kernels
            if (part == 1) {
                ctx->prof_pending.push_back({K_FACE_COEF, marks[0], marks[1]});
                ctx->prof_pool.push_back(marks[2]); ctx->prof_pool.push_back(marks[3]);
            } else if (part == 2) {
                ctx->prof_pool.push_back(marks[0]);
                ctx->prof_pending.push_back({K_UPWIND_PASS1, marks[1], marks[2]});
                ctx->prof_pending.push_back({-K_CORRECTOR - 100, marks[2], marks[3]});
            } else {
                ctx->prof_pending.push_back({K_FACE_COEF, marks[0], marks[1]});
                ctx->prof_pending.push_back({-K_UPWIND_PASS1 - 100, marks[1], marks[2]});    // negative id: the start event is the
                ctx->prof_pending.push_back({-K_CORRECTOR - 100, marks[2], marks[3]});       // previous record's stop event
            }
        }
        if (nl < 0) return fail(ctx, ICAR_ERR_CUDA, "icar_advect: cannot set up the TMA tensor maps / shared memory of the fused MPDATA kernel");
        ctx->launches += nl;
        CU(cudaGetLastError());
        if (part == 1) {
            if (ws != ctx->stream) { CU(cudaEventRecord(ctx->ev_winds, ws)); ctx->winds_pending = true; }
            return ICAR_OK;
        }
        for (int n = 0; n < ctx->nadv; ++n) {      // the output buffer becomes the field, the old field the next output
            float* old_field = (float*)ctx->vars[ctx->adv[n]];
            ctx->vars[ctx->adv[n]] = ctx->fout[n];
            ctx->fout[n] = old_field;
        }
        return ICAR_OK;
    }
    if (part == 1) {
        if (ws != ctx->stream) { CU(cudaEventRecord(ctx->ev_winds, ws)); ctx->winds_pending = true; }
        return ICAR_OK;
    }
    // adv_mpdata.f90:512-522: every scalar with vars_to_advect > 0 (mp_simple.f90:116-118)
    ProfScope ps_staged(ctx, K_STAGED_ADVECT3D, ctx->stream);
    for (int s = 0; s < ctx->nadv; ++s) {
        int var = ctx->adv[s];
        float* q = (float*)ctx->vars[var];
        float* res = nullptr;
        ctx->launches += staged_advect3d<float>(g, q, rho, jaco, dz, ctx->U, ctx->V, ctx->W, ctx->Wn, G, order,
                                                ctx->opt.flux_corrected_transport != 0, ctx->qa, ctx->qb,
                                                ctx->u2, ctx->v2, ctx->w2, &res, ctx->stream);
        // the field stays where it is (its buffer is part of the peer-mapped arena): copy the result home
        if (res != q) CU(cudaMemcpyAsync(q, res, cell_elems(ctx) * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CU(cudaGetLastError());
    return ICAR_OK;
}

extern "C" int icar_advect(icar_ctx* ctx, float dt) { return advect_impl(ctx, dt, 0, nullptr); }

// ============================================================================ mp

// mp() as step() needs it, without the reference's fixed ring geometry.  The only purpose of the
// mp(halo) / mp(subset) pair in time_step.f90:512-526 is to finish the cells a neighbour needs before
// the exchange starts; which cells those are depends on the decomposition:
//   phase 0 ("halo")   : the h rows/columns next to each ghost zone of the exchange axis
//   phase 1 ("subset") : every other cell of the physics tile
// Both phases together cover the tile exactly once, like process_halo + process_subdomain(subset).
int mp_phase(icar_ctx* ctx, float dt_in, int phase, int axis, int h);
int halo_send_from(icar_ctx* ctx, cudaStream_t producer);
int advect_impl(icar_ctx* ctx, float dt, int part, cudaStream_t winds_stream);

extern "C" int icar_mp(icar_ctx* ctx, float dt_in, int halo, int subset) {
    if (!ctx) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (ctx->opt.microphysics == 0) return ICAR_OK;
    int rc = need_uploaded(ctx, {ICAR_VAR_PRESSURE, ICAR_VAR_EXNER, ICAR_VAR_DENSITY, ICAR_VAR_DZ_MASS}, "icar_mp");
    if (rc) return rc;
    const icar_grid& gr = ctx->grid;
    const bool has_halo = halo >= 0, has_subset = subset >= 0;
    // mp_driver.f90:698-713
    if (ctx->last_model_time == -999.0)
        ctx->last_model_time = ctx->model_time - (double)std::max((float)ctx->opt.update_interval, dt_in);
    if (((ctx->model_time + (double)dt_in) - ctx->last_model_time) >= (double)ctx->opt.update_interval) {
        float mp_dt = (float)(ctx->model_time - ctx->last_model_time);
        if (!has_halo) ctx->last_model_time = ctx->model_time;
        int kte = gr.kte;
        if (ctx->opt.top_mp_level > 0) kte = std::min(kte, ctx->opt.top_mp_level);   // :723-725
        // process_subdomain's local precipitation arrays (:411-413, :587-595) are folded into the kernels
        if (has_halo || (!has_halo && !has_subset))
            if ((rc = join_aux(ctx))) return rc;          // earlier strips overlap these cells
        if (has_subset || (!has_halo && !has_subset)) {
            const int sub = has_subset ? subset : 0;                                     // :728-738 / :754-766
            const int t[1][4] = {{gr.its + sub, gr.ite - sub, gr.jts + sub, gr.jte - sub}};
            if ((rc = run_mp_tiles(ctx, mp_dt, t, 1, gr.kts, kte, ctx->stream))) return rc;
        }
        if (has_halo) {                                                                  // process_halo :609-658
            // The four strips are thin (latency-bound): they share ONE launch on an auxiliary stream
            // and are joined lazily, so the interior tile of mp(subset) overlaps them the way
            // time_step.f90:512-526 overlaps the halo exchange.
            CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
            const int t[4][4] = {{gr.its, gr.its + halo - 1, gr.jts, gr.jte},
                                 {gr.ite - halo + 1, gr.ite, gr.jts, gr.jte},
                                 {gr.its + halo, gr.ite - halo, gr.jts, gr.jts + halo - 1},
                                 {gr.its + halo, gr.ite - halo, gr.jte - halo + 1, gr.jte}};
            CU(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
            if ((rc = run_mp_tiles(ctx, mp_dt, t, 4, gr.kts, kte, ctx->aux))) return rc;
            CU(cudaEventRecord(ctx->ev_aux, ctx->aux));
            ctx->aux_pending = true;
        }
        CU(cudaGetLastError());
    }
    return ICAR_OK;
}

int mp_phase(icar_ctx* ctx, float dt_in, int phase, int axis, int h) {
    int rc = need_uploaded(ctx, {ICAR_VAR_PRESSURE, ICAR_VAR_EXNER, ICAR_VAR_DENSITY, ICAR_VAR_DZ_MASS}, "icar_physics_step");
    if (rc) return rc;
    if (ctx->opt.microphysics == 0) return ICAR_OK;
    const icar_grid& gr = ctx->grid;
    if (ctx->last_model_time == -999.0)                                                   // mp_driver.f90:698-700
        ctx->last_model_time = ctx->model_time - (double)std::max((float)ctx->opt.update_interval, dt_in);
    if (!(((ctx->model_time + (double)dt_in) - ctx->last_model_time) >= (double)ctx->opt.update_interval)) return ICAR_OK;
    const float mp_dt = (float)(ctx->model_time - ctx->last_model_time);
    if (phase == 1) ctx->last_model_time = ctx->model_time;                               // :711-713
    int kte = gr.kte;
    if (ctx->opt.top_mp_level > 0) kte = std::min(kte, ctx->opt.top_mp_level);
    // strips exist only next to a ghost zone (an interior edge of the decomposition)
    const bool lo = axis ? (gr.jms > gr.jds) : (gr.ims > gr.ids);
    const bool hi = axis ? (gr.jme < gr.jde) : (gr.ime < gr.ide);
    const int a0 = axis ? gr.jts : gr.its, a1 = axis ? gr.jte : gr.ite;
    const int in0 = lo ? a0 + h : a0, in1 = hi ? a1 - h : a1;
    if (phase == 0) {
        // the strips run on the auxiliary stream: the exchange waits for them only, and the interior tile
        // (phase 1, compute stream) runs next to both
        if ((rc = join_aux(ctx))) return rc;
        CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
        int t[2][4], n = 0;
        if (lo) { t[n][0] = axis ? gr.its : a0; t[n][1] = axis ? gr.ite : a0 + h - 1; t[n][2] = axis ? a0 : gr.jts; t[n][3] = axis ? a0 + h - 1 : gr.jte; ++n; }
        if (hi) { t[n][0] = axis ? gr.its : a1 - h + 1; t[n][1] = axis ? gr.ite : a1; t[n][2] = axis ? a1 - h + 1 : gr.jts; t[n][3] = axis ? a1 : gr.jte; ++n; }
        if ((rc = run_mp_tiles(ctx, mp_dt, t, n, gr.kts, kte, ctx->aux))) return rc;
        CU(cudaEventRecord(ctx->ev_aux, ctx->aux));
        ctx->aux_pending = true;
    } else {
        const int t[1][4] = {{axis ? gr.its : in0, axis ? gr.ite : in1, axis ? in0 : gr.jts, axis ? in1 : gr.jte}};
        if ((rc = run_mp_tiles(ctx, mp_dt, t, 1, gr.kts, kte, ctx->stream))) return rc;
    }
    return ICAR_OK;
}

// ============================================================================ halo exchange

extern "C" int icar_halo_plan(const icar_grid* grid, int image, int nimages, int plan[8]) {
    if (!grid || !plan || nimages < 1 || image < 1 || image > nimages) return ICAR_ERR_INVALID;
    const int h = grid->halo_size;
    const bool xcut = (grid->ims > grid->ids) || (grid->ime < grid->ide);
    const bool ycut = (grid->jms > grid->jds) || (grid->jme < grid->jde);
    if (xcut && ycut) return ICAR_ERR_INVALID;       // 2-D tilings are not supported (slabs only)
    const int axis = ycut ? 1 : 0;
    const int n = axis ? (grid->jme - grid->jms + 1) : (grid->ime - grid->ims + 1);
    if (h < 1 || n < 2 * h) return ICAR_ERR_INVALID;
    plan[0] = h;               // put_west / put_south
    plan[1] = n - 2 * h;       // put_east / put_north
    plan[2] = 0;               // retrieve_west_halo / retrieve_south_halo
    plan[3] = n - h;           // retrieve_east_halo / retrieve_north_halo
    plan[4] = image > 1 ? 1 : 0;
    plan[5] = image < nimages ? 1 : 0;
    plan[6] = axis;
    plan[7] = h;
    return ICAR_OK;
}

extern "C" int icar_comm_unique_id(char id[128]) {
    std::string err;
    if (!g_nccl.load(err)) { g_create_error = err; return ICAR_ERR_COMM; }
    ncclUniqueId uid;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (g_nccl.GetUniqueId(&uid) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return ICAR_ERR_COMM; }
    std::memcpy(id, &uid, 128);
    return ICAR_OK;
}

namespace {

// what an image tells the others about its arena (all-gathered at icar_comm_init)
struct ArenaInfo {
    cudaIpcMemHandle_t handle;
    unsigned long long slot_bytes;
    int nx, nz, ny, ld, nadv, device;
};

// Advected scalar n of `ctx` lives in home[n][b] with the same b on every image: all images make the same sequence of
// advect() calls, and only the fused advect() swaps (b flips for every scalar at once).
int current_buffer(const icar_ctx* ctx) { return ctx->vars[ctx->adv[0]] == (void*)ctx->home[0][0] ? 0 : 1; }

// The links of the exchange that starts now: this image's fields and flags, and for each neighbour its fields (in the
// buffers that hold them now), geometry and flags.
int halo_links(icar_ctx* ctx, HaloLinks* L) {
    int plan[8];
    if (icar_halo_plan(&ctx->grid, ctx->image, ctx->nimages, plan))
        return fail(ctx, ICAR_ERR_INVALID, "halo exchange: tile narrower than two halos, or not a slab decomposition");
    if (!ctx->linked) return fail(ctx, ICAR_ERR_COMM, "halo exchange: icar_comm_init / icar_comm_init_local was not called");
    std::memset(L, 0, sizeof(*L));
    L->nfields = ctx->nexch;
    L->axis = plan[6];
    L->h = plan[7];
    L->flags = (unsigned*)ctx->arena;
    const int b = current_buffer(ctx);
    int slot_of[ICAR_MAX_EXCHANGED];
    for (int s = 0; s < ctx->nexch; ++s) {
        const int n = (int)(std::find(ctx->adv, ctx->adv + ctx->nadv, ctx->exch[s]) - ctx->adv);
        if (n >= ctx->nadv || ctx->vars[ctx->exch[s]] != (void*)ctx->home[n][b])
            return fail(ctx, ICAR_ERR_INVALID, "halo exchange: an exchanged scalar is not in its arena buffer");
        slot_of[s] = 2 * n + b;
        L->src[s] = ctx->home[n][b];
    }
    for (int side = 0; side < 2; ++side) {
        HaloSide& S = L->side[side];
        S.present = plan[4 + side];
        if (!S.present) continue;
        const icar_ctx::Peer& P = ctx->peer[side];
        if (!P.base) return fail(ctx, ICAR_ERR_COMM, "halo exchange: neighbour not mapped");
        const int peer_extent = L->axis ? P.ny : P.nx;
        S.src0 = side == 0 ? plan[0] : plan[1];                  // put_west / put_south : put_east / put_north
        S.dst0 = side == 0 ? peer_extent - L->h : 0;             // its upper ghost cells : its lower ghost cells
        S.dst_ld = P.ld;
        S.dst_ny = P.ny;
        S.flags = (unsigned*)P.base;
        for (int s = 0; s < ctx->nexch; ++s)
            S.dst[s] = (float*)((char*)P.base + kArenaHeader + (size_t)slot_of[s] * P.slot_bytes);
    }
    return ICAR_OK;
}

int check_peer(icar_ctx* ctx, int side, const ArenaInfo& me, const ArenaInfo& other) {
    const bool ycut = (ctx->grid.jms > ctx->grid.jds) || (ctx->grid.jme < ctx->grid.jde);
    const bool ok = other.nz == me.nz && other.nadv == me.nadv && (ycut ? (other.nx == me.nx && other.ld == me.ld) : other.ny == me.ny);
    if (!ok) return fail(ctx, ICAR_ERR_COMM, "icar_comm_init: the neighbouring image has a different cross-section or scheme");
    ctx->peer[side].nx = other.nx; ctx->peer[side].ny = other.ny; ctx->peer[side].ld = other.ld;
    ctx->peer[side].slot_bytes = (size_t)other.slot_bytes;
    return ICAR_OK;
}

// every kernel of the multi-image step, loaded before the first exchange (preload_kernel, common.cuh)
int preload_step_kernels(icar_ctx* ctx) {
    const bool ok = halo_preload() && staged_preload() && fused_preload() && mp_simple_preload() && th_driver_preload();
    if (!ok) { (void)cudaGetLastError(); return fail(ctx, ICAR_ERR_CUDA, "icar_comm_init: cannot load the kernels of the step"); }
    return ICAR_OK;
}

ArenaInfo arena_info(const icar_ctx* ctx) {
    ArenaInfo a{};
    a.slot_bytes = ctx->slot_bytes;
    a.nx = ctx->g.nx; a.nz = ctx->g.nz; a.ny = ctx->g.ny; a.ld = ctx->g.ld; a.nadv = ctx->nadv; a.device = ctx->device;
    return a;
}

}  // namespace

// One process per image: NCCL communicator (co_min of update_dt) + the neighbours' arenas mapped with CUDA IPC.
extern "C" int icar_comm_init(icar_ctx* ctx, int image, int nimages, const char id[128]) {
    if (!ctx || !id || image < 1 || image > nimages) return fail(ctx, ICAR_ERR_INVALID, "icar_comm_init: bad argument");
    CU(cudaSetDevice(ctx->device));
    ctx->image = image;
    ctx->nimages = nimages;
    if (nimages == 1) return ICAR_OK;
    { int rc = preload_step_kernels(ctx); if (rc) return rc; }
    std::string err;
    if (!g_nccl.load(err)) return fail(ctx, ICAR_ERR_COMM, err);
    ncclUniqueId uid;
    std::memcpy(&uid, id, 128);
    NC(g_nccl.CommInitRank(&ctx->comm, nimages, uid, image - 1));
    // all-gather the arena descriptions
    ArenaInfo mine = arena_info(ctx);
    CU(cudaIpcGetMemHandle(&mine.handle, ctx->arena));
    std::vector<ArenaInfo> all(nimages);
    char* d = nullptr;
    CU(cudaMalloc((void**)&d, sizeof(ArenaInfo) * (size_t)(nimages + 1)));
    cudaError_t e = cudaMemcpyAsync(d, &mine, sizeof(ArenaInfo), cudaMemcpyHostToDevice, ctx->stream);
    ncclResult_t nr = ncclSuccess;
    if (e == cudaSuccess) nr = g_nccl.AllGather(d, d + sizeof(ArenaInfo), sizeof(ArenaInfo), ncclChar, ctx->comm, ctx->stream);
    if (e == cudaSuccess && nr == ncclSuccess)
        e = cudaMemcpyAsync(all.data(), d + sizeof(ArenaInfo), sizeof(ArenaInfo) * (size_t)nimages, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && nr == ncclSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    NC(nr);
    CU(e);
    for (int side = 0; side < 2; ++side) {
        const int other = side == 0 ? image - 2 : image;          // 0-based rank of the neighbour
        if (other < 0 || other >= nimages) continue;
        int rc = check_peer(ctx, side, mine, all[other]);
        if (rc) return rc;
        void* base = nullptr;
        const cudaError_t oe = cudaIpcOpenMemHandle(&base, all[other].handle, cudaIpcMemLazyEnablePeerAccess);
        if (oe != cudaSuccess)
            return fail(ctx, ICAR_ERR_COMM, std::string("icar_comm_init: cannot map the neighbouring image's memory (cudaIpcOpenMemHandle: ") +
                                                cudaGetErrorString(oe) + "); the images of one run must be on GPUs of one box with peer access");
        ctx->peer[side].base = base;
        ctx->peer[side].ipc = true;
    }
    ctx->linked = true;
    return ICAR_OK;
}

// All images in ONE process (a test harness, or one host thread driving several GPUs): ctxs[i] is image i + 1.  No NCCL
// communicator: icar_compute_dt then returns the image's own time step and the caller takes the minimum (co_min).
// The caller must queue the steps of all images before it waits for any of them (an image's kernels wait on the device
// for its neighbours' data).
extern "C" int icar_comm_init_local(icar_ctx** ctxs, int nimages) {
    if (!ctxs || nimages < 1) return ICAR_ERR_INVALID;
    for (int i = 0; i < nimages; ++i) if (!ctxs[i]) return ICAR_ERR_INVALID;
    for (int i = 0; i < nimages; ++i) {
        icar_ctx* ctx = ctxs[i];
        CU(cudaSetDevice(ctx->device));
        ctx->image = i + 1;
        ctx->nimages = nimages;
        if (nimages == 1) continue;
        { int rc = preload_step_kernels(ctx); if (rc) return rc; }
        const ArenaInfo mine = arena_info(ctx);
        for (int side = 0; side < 2; ++side) {
            const int other = side == 0 ? i - 1 : i + 1;
            if (other < 0 || other >= nimages) continue;
            int rc = check_peer(ctx, side, mine, arena_info(ctxs[other]));
            if (rc) return rc;
            if (ctxs[other]->device != ctx->device) {
                const cudaError_t pe = cudaDeviceEnablePeerAccess(ctxs[other]->device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(ctx, ICAR_ERR_COMM, std::string("icar_comm_init_local: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe));
                (void)cudaGetLastError();
            }
            ctx->peer[side].base = ctxs[other]->arena;
            ctx->peer[side].ipc = false;
        }
        ctx->linked = true;
    }
    return ICAR_OK;
}

// halo_send with the stream that produced the cells to be sent: the exchange is ordered after the work queued on
// `producer` only.  `producer` must itself be ordered after everything that read this image's ghost cells (the previous
// advect): the auxiliary stream forks from the compute stream, so it is.
int halo_send_from(icar_ctx* ctx, cudaStream_t producer);

extern "C" int icar_halo_send(icar_ctx* ctx) {
    if (!ctx) return ICAR_ERR_INVALID;
    if (ctx->nimages == 1) return ICAR_OK;
    CU(cudaSetDevice(ctx->device));
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    return halo_send_from(ctx, ctx->stream);
}

int halo_send_from(icar_ctx* ctx, cudaStream_t producer) {
    if (ctx->exchange_in_flight) return fail(ctx, ICAR_ERR_INVALID, "icar_halo_send: the previous exchange has not been retrieved");
    HaloLinks L;
    int rc = halo_links(ctx, &L);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev_packed, producer));
    CU(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_packed, 0));
    {
        ProfScope ps(ctx, K_HALO_PUSH, ctx->comm_stream);
        ctx->launches += halo_begin_launch(L, ctx->comm_stream);
        ctx->launches += halo_push_launch(ctx->g, L, ctx->comm_stream);
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
    ctx->exchange_in_flight = true;
    return ICAR_OK;
}

extern "C" int icar_halo_retrieve(icar_ctx* ctx) {
    if (!ctx) return ICAR_ERR_INVALID;
    if (ctx->nimages == 1) return ICAR_OK;
    if (!ctx->exchange_in_flight) return fail(ctx, ICAR_ERR_INVALID, "icar_halo_retrieve without icar_halo_send");
    CU(cudaSetDevice(ctx->device));
    HaloLinks L;
    int rc = halo_links(ctx, &L);
    if (rc) return rc;
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));     // this image's own puts are out
    {
        ProfScope ps(ctx, K_HALO_WAIT, ctx->stream);           // `sync images(neighbors)`, exchangeable_obj.f90:151
        ctx->launches += halo_wait_launch(L, ctx->stream);
    }
    CU(cudaGetLastError());
    ctx->exchange_in_flight = false;
    return ICAR_OK;
}

extern "C" int icar_halo_exchange(icar_ctx* ctx) {
    int rc = icar_halo_send(ctx);
    return rc ? rc : icar_halo_retrieve(ctx);
}

// ============================================================================ step

extern "C" int icar_physics_step(icar_ctx* ctx, float dt) {
    if (!ctx) return ICAR_ERR_INVALID;
    int rc;
    const int h = ctx->grid.halo_size;
    if (dt > 1e-3f) {                                             // time_step.f90:484
        if (ctx->nimages == 1) {
            // the wind-only part of advect() goes to the auxiliary stream first: HBM-bound kernels next to the
            // issue-bound microphysics kernel (they share no data)
            if ((rc = join_aux(ctx))) return rc;
            CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
            CU(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
            if ((rc = advect_impl(ctx, dt, 1, ctx->aux))) return rc;
            // nothing to exchange: mp(halo=1) + mp(subset=1) of :512,:523 is one whole-tile call
            if ((rc = icar_mp(ctx, dt, -1, -1))) return rc;
            if ((rc = advect_impl(ctx, dt, 2, nullptr))) return rc;   // :529
            ctx->model_time += (double)dt;                            // :547
            ++ctx->eager_steps;
            return ICAR_OK;
        } else {
            int plan[8];
            if (icar_halo_plan(&ctx->grid, ctx->image, ctx->nimages, plan)) return fail(ctx, ICAR_ERR_INVALID, "icar_physics_step: not a slab decomposition");
            bool due = ctx->opt.microphysics != 0;
            if (due) {                                                // mp_driver.f90:698-703
                const double last = ctx->last_model_time == -999.0
                    ? ctx->model_time - (double)std::max((float)ctx->opt.update_interval, dt) : ctx->last_model_time;
                due = ((ctx->model_time + (double)dt) - last) >= (double)ctx->opt.update_interval;
            }
            if (!due) {
                // no strips this step: the auxiliary stream carries the halo pack alone, which still has to
                // start after everything queued on the compute stream
                if ((rc = join_aux(ctx))) return rc;
                CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
                CU(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
            }
            if ((rc = mp_phase(ctx, dt, 0, plan[6], h))) return rc;   // :512  the cells the neighbours need (aux stream)
            if ((rc = halo_send_from(ctx, ctx->aux))) return rc;      // :515  ordered after the strips only
            if ((rc = advect_impl(ctx, dt, 1, ctx->aux))) return rc;  // the wind-only part of advect() needs no halo: behind the strips
            if ((rc = mp_phase(ctx, dt, 1, plan[6], h))) return rc;   // :523  everything else, overlapping the exchange
            if ((rc = icar_halo_retrieve(ctx))) return rc;            // :526
            if ((rc = advect_impl(ctx, dt, 2, nullptr))) return rc;   // :529
            ctx->model_time += (double)dt;                            // :547
            ++ctx->eager_steps;
            return ICAR_OK;
        }
        if ((rc = icar_advect(ctx, dt))) return rc;               // :529
        ++ctx->eager_steps;
    }
    ctx->model_time += (double)dt;                                // :547
    return ICAR_OK;
}

// ---------------------------------------------------------------------------- several steps, as a CUDA graph

namespace {

// What mp() would decide at `model_time` (mp_driver.f90:698-713), without running anything.
void mp_schedule(const icar_ctx* ctx, double model_time, double last, float dt, bool* runs, float* mp_dt, double* new_last) {
    *runs = false; *mp_dt = 0.f; *new_last = last;
    if (ctx->opt.microphysics == 0) return;
    if (((model_time + (double)dt) - last) >= (double)ctx->opt.update_interval) {
        *runs = true;
        *mp_dt = (float)(model_time - last);
        *new_last = model_time;
    }
}

void drop_step_graph(icar_ctx* ctx) {
    if (ctx->sgraph.exec) cudaGraphExecDestroy(ctx->sgraph.exec);
    ctx->sgraph.exec = nullptr;
}

// Runs steps A and B of a pair through the captured graph when that is valid; *ran says whether it did.
int graph_pair(icar_ctx* ctx, float dt, bool* ran) {
    *ran = false;
    icar_ctx::StepGraph& sg = ctx->sgraph;
    // the eager path has to have run first: it does the lazy allocations and sets the mp_dt lag up
    if (sg.unsupported || ctx->prof_on || ctx->eager_steps < 2 || ctx->last_model_time == -999.0 ||
        ctx->aux_pending || ctx->winds_pending || ctx->exchange_in_flight)
        return ICAR_OK;
    bool runs[2]; float mp_dt[2]; double last = ctx->last_model_time, t = ctx->model_time;
    mp_schedule(ctx, t, last, dt, &runs[0], &mp_dt[0], &last);
    mp_schedule(ctx, t + (double)dt, last, dt, &runs[1], &mp_dt[1], &last);
    const double t_end = (t + (double)dt) + (double)dt;
    bool match = sg.exec && sg.dt == dt && std::memcmp(sg.vars, ctx->vars, sizeof(sg.vars)) == 0;
    for (int a = 0; a < 2 && match; ++a) match = sg.runs[a] == runs[a] && (!runs[a] || sg.mp_dt[a] == mp_dt[a]);
    if (!match) {
        drop_step_graph(ctx);
        // Capture queues nothing on the GPU, but the two calls advance the host-side state (clock, launch
        // counter, buffer roles): snapshot, capture, and put it back.
        const double t0 = ctx->model_time, last0 = ctx->last_model_time;
        const long long launches0 = ctx->launches, eager0 = ctx->eager_steps;
        void* vars0[ICAR_VAR_COUNT]; std::memcpy(vars0, ctx->vars, sizeof(vars0));
        float* fout0[ICAR_MAX_SCALARS]; std::memcpy(fout0, ctx->fout, sizeof(fout0));
        float *qa0 = ctx->qa, *qb0 = ctx->qb;
        CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
        int rc = icar_physics_step(ctx, dt);
        if (!rc) rc = icar_physics_step(ctx, dt);
        if (!rc) rc = join_aux(ctx);
        if (!rc) rc = join_winds(ctx);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);      // always: leaves capture mode
        const long long pair_launches = ctx->launches - launches0;
        const bool same_buffers = std::memcmp(vars0, ctx->vars, sizeof(vars0)) == 0 &&
                                  std::memcmp(fout0, ctx->fout, sizeof(fout0)) == 0 && qa0 == ctx->qa && qb0 == ctx->qb;
        ctx->model_time = t0; ctx->last_model_time = last0; ctx->launches = launches0; ctx->eager_steps = eager0;
        std::memcpy(ctx->vars, vars0, sizeof(vars0)); std::memcpy(ctx->fout, fout0, sizeof(fout0));
        ctx->qa = qa0; ctx->qb = qb0;
        ctx->aux_pending = false; ctx->winds_pending = false; ctx->exchange_in_flight = false;
        if (rc || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            sg.unsupported = true;
            if (rc) return rc;
            return fail(ctx, ICAR_ERR_CUDA, std::string("icar_physics_steps: stream capture failed: ") + cudaGetErrorString(ce));
        }
        if (!same_buffers) {             // the buffer roles have a longer period than two steps: no graph
            cudaGraphDestroy(graph);
            sg.unsupported = true;
            return ICAR_OK;
        }
        const cudaError_t ie = cudaGraphInstantiate(&sg.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
            sg.exec = nullptr; sg.unsupported = true;
            return fail(ctx, ICAR_ERR_CUDA, std::string("icar_physics_steps: cudaGraphInstantiate: ") + cudaGetErrorString(ie));
        }
        sg.dt = dt;
        for (int a = 0; a < 2; ++a) { sg.runs[a] = runs[a]; sg.mp_dt[a] = mp_dt[a]; }
        std::memcpy(sg.vars, ctx->vars, sizeof(sg.vars));
        sg.launches = pair_launches;
    }
    CU(cudaGraphLaunch(sg.exec, ctx->stream));
    ctx->model_time = t_end;
    ctx->last_model_time = last;
    ctx->launches += sg.launches;
    ctx->eager_steps += 2;
    *ran = true;
    return ICAR_OK;
}

}  // namespace

extern "C" int icar_physics_steps(icar_ctx* ctx, float dt, int nsteps, int use_graph) {
    if (!ctx || nsteps < 0) return fail(ctx, ICAR_ERR_INVALID, "icar_physics_steps: bad argument");
    CU(cudaSetDevice(ctx->device));
    int done = 0, rc;
    while (done < nsteps) {
        if (use_graph && nsteps - done >= 2 && dt > 1e-3f) {
            bool ran = false;
            if ((rc = graph_pair(ctx, dt, &ran))) return rc;
            if (ran) { done += 2; continue; }
        }
        if ((rc = icar_physics_step(ctx, dt))) return rc;
        ++done;
    }
    return ICAR_OK;
}

// ============================================================================ host-pointer leaf calls

namespace {

template <class R>
int leaf_set_winds(icar_ctx* ctx, const R* u, const R* v, const R* w, const R* ju, const R* jv, const R* jw,
                   const R* rho, R dt, R dx, int scheme, size_t slot0) {
    const Geom& g = ctx->g;
    const size_t es = sizeof(R);
    void* b[10];
    for (int i = 0; i < 10; ++i) { int rc = get_scratch(ctx, slot0 + i, es, &b[i]); if (rc) return rc; }
    const size_t rows = (size_t)g.nz * g.ny;
    int rc;
    if ((rc = copy_in(ctx, b[0], u, g.nx + 1, g.ldu, rows, es))) return rc;
    if ((rc = copy_in(ctx, b[1], v, g.nx, g.ld, (size_t)g.nz * (g.ny + 1), es))) return rc;
    if ((rc = copy_in(ctx, b[2], w, g.nx, g.ld, rows, es))) return rc;
    if ((rc = copy_in(ctx, b[3], ju, g.nx + 1, g.ldu, rows, es))) return rc;
    if ((rc = copy_in(ctx, b[4], jv, g.nx, g.ld, (size_t)g.nz * (g.ny + 1), es))) return rc;
    if ((rc = copy_in(ctx, b[5], jw, g.nx, g.ld, rows, es))) return rc;
    if (rho && (rc = copy_in(ctx, b[6], rho, g.nx, g.ld, rows, es))) return rc;
    ctx->launches += staged_wind_normalise<R>(g, (R*)b[0], (R*)b[1], (R*)b[2], (R*)b[3], (R*)b[4], (R*)b[5],
                                              rho ? (R*)b[6] : nullptr, dt, dx, scheme, (R*)b[7], (R*)b[8],
                                              (R*)b[9], ctx->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    return ICAR_OK;
}

template <class R>
int leaf_advect3d(icar_ctx* ctx, R* q, const R* rho, const R* jaco, const R* dz, int order, bool fct,
                  size_t slot0) {
    const Geom& g = ctx->g;
    const size_t es = sizeof(R);
    void* b[22];
    for (int i = 0; i < 22; ++i) { int rc = get_scratch(ctx, slot0 + i, es, &b[i]); if (rc) return rc; }
    R *U = (R*)b[7], *V = (R*)b[8], *W = (R*)b[9];
    R *dq = (R*)b[10], *drho = (R*)b[11], *djaco = (R*)b[12], *ddz = (R*)b[13], *Wn = (R*)b[14], *G = (R*)b[15];
    R *qa = (R*)b[16], *qb = (R*)b[17], *u2 = (R*)b[18], *v2 = (R*)b[19], *w2 = (R*)b[20];
    const size_t rows = (size_t)g.nz * g.ny;
    int rc;
    if ((rc = copy_in(ctx, dq, q, g.nx, g.ld, rows, es))) return rc;
    if (rho && (rc = copy_in(ctx, drho, rho, g.nx, g.ld, rows, es))) return rc;
    if ((rc = copy_in(ctx, djaco, jaco, g.nx, g.ld, rows, es))) return rc;
    if ((rc = copy_in(ctx, ddz, dz, g.nx, g.ld, rows, es))) return rc;
    const R* Gp = djaco;
    if (order > 1) {
        ctx->launches += staged_wn_G<R>(g, W, ddz, djaco, rho ? drho : nullptr, Wn, rho ? G : nullptr, ctx->stream);
        if (rho) Gp = G;
    }
    R* res = nullptr;
    ctx->launches += staged_advect3d<R>(g, dq, rho ? drho : nullptr, djaco, ddz, U, V, W, Wn, Gp, order, fct, qa,
                                        qb, u2, v2, w2, &res, ctx->stream);
    CU(cudaGetLastError());
    if ((rc = copy_out(ctx, q, res, g.nx, g.ld, rows, es))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return ICAR_OK;
}

}  // namespace

// scratch slots: 0..21 float leaf calls, 32..53 double leaf calls, 64.. mp_simple leaf call
extern "C" int icar_mpdata_set_winds(icar_ctx* ctx, const float* u, const float* v, const float* w,
                                     const float* ju, const float* jv, const float* jw, const float* rho,
                                     float dt, float dx, int scheme) {
    if (!ctx || !u || !v || !w || !ju || !jv || !jw) return fail(ctx, ICAR_ERR_INVALID, "icar_mpdata_set_winds: null array");
    CU(cudaSetDevice(ctx->device));
    return leaf_set_winds<float>(ctx, u, v, w, ju, jv, jw, rho, dt, dx, scheme, 0);
}

extern "C" int icar_mpdata_advect3d(icar_ctx* ctx, float* q, const float* rho, int ims, int ime, int kms,
                                    int kme, int jms, int jme, const float* jaco, const float* dz, int order,
                                    int fct) {
    if (!ctx || !q || !jaco || !dz) return fail(ctx, ICAR_ERR_INVALID, "icar_mpdata_advect3d: null array");
    if (!check_dims(ctx, ims, ime, kms, kme, jms, jme)) return fail(ctx, ICAR_ERR_INVALID, "icar_mpdata_advect3d: bounds do not match the context's tile");
    if (order < 1) return ICAR_OK;      // `do iord=1,mpdata_order` runs zero times
    if (ctx->scratch.size() < 10 || !ctx->scratch[9]) return fail(ctx, ICAR_ERR_INVALID, "icar_mpdata_advect3d before icar_mpdata_set_winds");
    CU(cudaSetDevice(ctx->device));
    return leaf_advect3d<float>(ctx, q, rho, jaco, dz, order, fct != 0, 0);
}

extern "C" int icar_mpdata_set_winds_f64(icar_ctx* ctx, const double* u, const double* v, const double* w,
                                         const double* ju, const double* jv, const double* jw,
                                         const double* rho, double dt, double dx, int scheme) {
    if (!ctx || !u || !v || !w || !ju || !jv || !jw) return fail(ctx, ICAR_ERR_INVALID, "icar_mpdata_set_winds_f64: null array");
    CU(cudaSetDevice(ctx->device));
    int rc = leaf_set_winds<double>(ctx, u, v, w, ju, jv, jw, rho, dt, dx, scheme, 32);
    if (!rc) ctx->winds_f64_valid = true;
    return rc;
}

extern "C" int icar_mpdata_advect3d_f64(icar_ctx* ctx, double* q, const double* rho, int ims, int ime,
                                        int kms, int kme, int jms, int jme, const double* jaco,
                                        const double* dz, int order, int fct) {
    if (!ctx || !q || !jaco || !dz) return fail(ctx, ICAR_ERR_INVALID, "icar_mpdata_advect3d_f64: null array");
    if (!check_dims(ctx, ims, ime, kms, kme, jms, jme)) return fail(ctx, ICAR_ERR_INVALID, "icar_mpdata_advect3d_f64: bounds do not match the context's tile");
    if (order < 1) return ICAR_OK;
    if (!ctx->winds_f64_valid) return fail(ctx, ICAR_ERR_INVALID, "icar_mpdata_advect3d_f64 before icar_mpdata_set_winds_f64");
    CU(cudaSetDevice(ctx->device));
    return leaf_advect3d<double>(ctx, q, rho, jaco, dz, order, fct != 0, 32);
}

extern "C" int icar_mp_simple_driver(icar_ctx* ctx, const float* pressure, float* th, const float* pii,
                                     const float* rho, float* qv, float* qc, float* qr, float* qs,
                                     float* rain, float* snow, float dt, const float* dz, int ims, int ime,
                                     int jms, int jme, int kms, int kme, int its, int ite, int jts, int jte,
                                     int kts, int kte) {
    if (!ctx || !pressure || !th || !pii || !rho || !qv || !qc || !qr || !qs || !rain || !snow || !dz)
        return fail(ctx, ICAR_ERR_INVALID, "icar_mp_simple_driver: null array");
    if (!check_dims(ctx, ims, ime, kms, kme, jms, jme)) return fail(ctx, ICAR_ERR_INVALID, "icar_mp_simple_driver: bounds do not match the context's tile");
    if (its < ims || ite > ime || jts < jms || jte > jme || kts < kms || kte > kme)
        return fail(ctx, ICAR_ERR_INVALID, "icar_mp_simple_driver: tile outside memory bounds");
    CU(cudaSetDevice(ctx->device));
    const Geom& g = ctx->g;
    void* b[11];
    for (int i = 0; i < 11; ++i) { int rc = get_scratch(ctx, 64 + i, sizeof(float), &b[i]); if (rc) return rc; }
    const size_t rows = (size_t)g.nz * g.ny;
    const float* in3[9] = {pressure, th, pii, rho, qv, qc, qr, qs, dz};
    int rc;
    for (int i = 0; i < 9; ++i)
        if ((rc = copy_in(ctx, b[i], in3[i], g.nx, g.ld, rows, sizeof(float)))) return rc;
    if ((rc = copy_in(ctx, b[9], rain, g.nx, ctx->ld2, g.ny, sizeof(float)))) return rc;
    if ((rc = copy_in(ctx, b[10], snow, g.nx, ctx->ld2, g.ny, sizeof(float)))) return rc;
    float c2s = std::exp(-1.0f * (1.0f / 2000.0f) * dt);
    float c2r = std::exp(-1.0f * (1.0f / 500.0f) * dt);
    MpTiles tiles;
    tiles.n = 1;
    tiles.its[0] = its - ims; tiles.ite[0] = ite - ims; tiles.jts[0] = jts - jms; tiles.jte[0] = jte - jms;
    tiles.along_y[0] = 0;
    {
        int nl = mp_simple_launch(g, (float*)b[0], (float*)b[1], (float*)b[2], (float*)b[3], (float*)b[4],
                                  (float*)b[5], (float*)b[6], (float*)b[7], (float*)b[9], (float*)b[10], ctx->ld2, dt,
                                  (float*)b[8], tiles, kts - kms, kte - kms, c2r, c2s, ctx->diag, ctx->stream);
        if (nl < 0) return fail(ctx, ICAR_ERR_INVALID, "mp_simple: more than 80 levels are not built");
        ctx->launches += nl;
    }
    CU(cudaGetLastError());
    float* out3[5] = {th, qv, qc, qr, qs};
    const int idx3[5] = {1, 4, 5, 6, 7};
    for (int i = 0; i < 5; ++i)
        if ((rc = copy_out(ctx, out3[i], b[idx3[i]], g.nx, g.ld, rows, sizeof(float)))) return rc;
    if ((rc = copy_out(ctx, rain, b[9], g.nx, ctx->ld2, g.ny, sizeof(float)))) return rc;
    if ((rc = copy_out(ctx, snow, b[10], g.nx, ctx->ld2, g.ny, sizeof(float)))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return ICAR_OK;
}

// ============================================================================ host-domain step

namespace {
struct HostField { int var; const void* in; void* out; };

void host_fields(const icar_host_domain* d, std::vector<HostField>& f) {
    f = {
        {ICAR_VAR_U, d->u, nullptr}, {ICAR_VAR_V, d->v, nullptr}, {ICAR_VAR_W, d->w, nullptr},
        {ICAR_VAR_PRESSURE, d->pressure, nullptr}, {ICAR_VAR_EXNER, d->exner, nullptr},
        {ICAR_VAR_DENSITY, d->density, nullptr},
        {ICAR_VAR_POTENTIAL_TEMPERATURE, d->potential_temperature, d->potential_temperature},
        {ICAR_VAR_WATER_VAPOR, d->water_vapor, d->water_vapor},
        {ICAR_VAR_CLOUD_WATER, d->cloud_water, d->cloud_water},
        {ICAR_VAR_RAIN_IN_AIR, d->rain_in_air, d->rain_in_air},
        {ICAR_VAR_SNOW_IN_AIR, d->snow_in_air, d->snow_in_air},
        {ICAR_VAR_PRECIPITATION, d->accumulated_precipitation, d->accumulated_precipitation},
        {ICAR_VAR_SNOWFALL, d->accumulated_snowfall, d->accumulated_snowfall},
        {ICAR_VAR_CLOUD_ICE, d->cloud_ice, d->cloud_ice},
        {ICAR_VAR_ICE_NUMBER, d->ice_number, d->ice_number},
        {ICAR_VAR_RAIN_NUMBER, d->rain_number, d->rain_number},
        {ICAR_VAR_GRAUPEL_IN_AIR, d->graupel_in_air, d->graupel_in_air},
        {ICAR_VAR_GRAUPEL, d->accumulated_graupel, d->accumulated_graupel},
    };
}
}  // namespace

extern "C" int icar_physics_step_host(icar_ctx* ctx, float dt, const icar_host_domain* dom) {
    if (!ctx || !dom) return fail(ctx, ICAR_ERR_INVALID, "icar_physics_step_host: null argument");
    CU(cudaSetDevice(ctx->device));
    std::vector<HostField> f;
    host_fields(dom, f);
    int rc;
    for (const HostField& h : f) {
        if (!h.in) continue;
        int width, pitch; size_t rows;
        var_geometry(ctx, h.var, &width, &pitch, &rows);
        if ((rc = copy_in(ctx, ctx->vars[h.var], h.in, width, pitch, rows, elem_size(h.var)))) return rc;
        ctx->uploaded[h.var] = true;
    }
    if ((rc = icar_physics_step(ctx, dt))) return rc;
    for (const HostField& h : f) {
        if (!h.out) continue;
        int width, pitch; size_t rows;
        var_geometry(ctx, h.var, &width, &pitch, &rows);
        if ((rc = copy_out(ctx, h.out, ctx->vars[h.var], width, pitch, rows, elem_size(h.var)))) return rc;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return ICAR_OK;
}

// diagnostic_update + the physics step on host arrays: exner and density are not uploaded but taken on the
// device from the pressure and potential temperature just uploaded (time_step.f90:85-101), which is what
// step()'s loop does before the physics and saves two of the thirteen per-step input fields on the PCIe side.
extern "C" int icar_step_host(icar_ctx* ctx, float dt, const icar_host_domain* dom) {
    if (!ctx || !dom) return fail(ctx, ICAR_ERR_INVALID, "icar_step_host: null argument");
    if (dom->exner || dom->density)
        return fail(ctx, ICAR_ERR_INVALID, "icar_step_host: exner and density are derived on the device; pass them as NULL "
                                           "(icar_physics_step_host takes them from the host)");
    CU(cudaSetDevice(ctx->device));
    std::vector<HostField> f;
    host_fields(dom, f);
    int rc;
    for (const HostField& h : f) {
        if (!h.in) continue;
        int width, pitch; size_t rows;
        var_geometry(ctx, h.var, &width, &pitch, &rows);
        if ((rc = copy_in(ctx, ctx->vars[h.var], h.in, width, pitch, rows, elem_size(h.var)))) return rc;
        ctx->uploaded[h.var] = true;
    }
    if ((rc = icar_diagnostic_update(ctx))) return rc;
    if ((rc = icar_physics_step(ctx, dt))) return rc;
    for (const HostField& h : f) {
        if (!h.out) continue;
        int width, pitch; size_t rows;
        var_geometry(ctx, h.var, &width, &pitch, &rows);
        if ((rc = copy_out(ctx, h.out, ctx->vars[h.var], width, pitch, rows, elem_size(h.var)))) return rc;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return ICAR_OK;
}

// The host step with separate input and output arrays and a row range for the outputs: what a caller needs
// to pipeline one model step over several contexts that each hold a band of rows (icar_b200/pipeline.py).
// `in`: every non-NULL member is uploaded (the whole tile).  `out`: every non-NULL output member receives the
// tile's rows row0..row1 (0-based y index, inclusive); the pointer addresses the tile's row 0.
namespace {
int host_io_impl(icar_ctx* ctx, float dt, const icar_host_domain* in, const icar_host_domain* out, int row0, int row1,
                 int derive_diagnostics, bool wait);
}

extern "C" int icar_physics_step_host_io(icar_ctx* ctx, float dt, const icar_host_domain* in, const icar_host_domain* out,
                                         int row0, int row1, int derive_diagnostics) {
    return host_io_impl(ctx, dt, in, out, row0, row1, derive_diagnostics, true);
}

extern "C" int icar_physics_step_host_io_async(icar_ctx* ctx, float dt, const icar_host_domain* in,
                                               const icar_host_domain* out, int row0, int row1, int derive_diagnostics) {
    return host_io_impl(ctx, dt, in, out, row0, row1, derive_diagnostics, false);
}

namespace {
int host_io_impl(icar_ctx* ctx, float dt, const icar_host_domain* in, const icar_host_domain* out, int row0, int row1,
                 int derive_diagnostics, bool wait) {
    if (!ctx || !in || !out) return fail(ctx, ICAR_ERR_INVALID, "icar_physics_step_host_io: null argument");
    if (row0 < 0 || row1 >= ctx->g.ny || row1 < row0) return fail(ctx, ICAR_ERR_INVALID, "icar_physics_step_host_io: bad row range");
    if (derive_diagnostics && (in->exner || in->density))
        return fail(ctx, ICAR_ERR_INVALID, "icar_physics_step_host_io: exner and density are derived on the device; pass them as NULL");
    CU(cudaSetDevice(ctx->device));
    std::vector<HostField> fi, fo;
    host_fields(in, fi);
    host_fields(out, fo);
    int rc;
    for (const HostField& h : fi) {
        if (!h.in) continue;
        int width, pitch; size_t rows;
        var_geometry(ctx, h.var, &width, &pitch, &rows);
        if ((rc = copy_in(ctx, ctx->vars[h.var], h.in, width, pitch, rows, elem_size(h.var)))) return rc;
        ctx->uploaded[h.var] = true;
    }
    if (derive_diagnostics && (rc = icar_diagnostic_update(ctx))) return rc;
    if ((rc = icar_physics_step(ctx, dt))) return rc;
    for (const HostField& h : fo) {
        if (!h.out) continue;
        int width, pitch; size_t rows;
        var_geometry(ctx, h.var, &width, &pitch, &rows);
        const size_t es = elem_size(h.var);
        const size_t lines_per_row = (var_shape(h.var) == D2) ? 1 : (size_t)ctx->g.nz;       // x-lines per y index
        const size_t first = lines_per_row * (size_t)row0, count = lines_per_row * (size_t)(row1 - row0 + 1);
        if ((rc = copy_out(ctx, (char*)h.out + first * (size_t)width * es,
                           (const char*)ctx->vars[h.var] + first * (size_t)pitch * es, width, pitch, count, es))) return rc;
    }
    if (wait) CU(cudaStreamSynchronize(ctx->stream));
    return ICAR_OK;
}
}  // namespace

extern "C" int icar_host_register(icar_ctx* ctx, const icar_host_domain* dom) {
    if (!ctx || !dom) return fail(ctx, ICAR_ERR_INVALID, "icar_host_register: null argument");
    CU(cudaSetDevice(ctx->device));
    std::vector<HostField> f;
    host_fields(dom, f);
    for (const HostField& h : f) {
        if (!h.in) continue;
        int width, pitch; size_t rows;
        var_geometry(ctx, h.var, &width, &pitch, &rows);
        CU(cudaHostRegister(const_cast<void*>(h.in), (size_t)width * rows * elem_size(h.var), cudaHostRegisterDefault));
    }
    return ICAR_OK;
}

extern "C" int icar_host_unregister(icar_ctx* ctx, const icar_host_domain* dom) {
    if (!ctx || !dom) return fail(ctx, ICAR_ERR_INVALID, "icar_host_unregister: null argument");
    CU(cudaSetDevice(ctx->device));
    std::vector<HostField> f;
    host_fields(dom, f);
    for (const HostField& h : f) {
        if (!h.in) continue;
        CU(cudaHostUnregister(const_cast<void*>(h.in)));
    }
    return ICAR_OK;
}

// ============================================================================ Thompson (mp = 1)

extern "C" void icar_default_mp_options(icar_mp_options* o) {
    if (!o) return;
    // options_obj.f90:1259-1280
    o->Nt_c = 100.e6f; o->TNO = 5.0f; o->am_s = 0.069f; o->rho_g = 500.0f;
    o->av_s = 40.0f; o->bv_s = 0.55f; o->fv_s = 100.0f; o->av_i = 1847.5f;
    o->av_g = 442.0f; o->bv_g = 0.89f;
    o->Ef_si = 0.05f; o->Ef_rs = 0.95f; o->Ef_rg = 0.75f; o->Ef_ri = 0.95f;
    o->C_cubes = 0.5f; o->C_sqrd = 0.3f; o->mu_r = 0.f; o->t_adjust = 0.f;
    o->Ef_rw_l = 0; o->Ef_sw_l = 0;
}

extern "C" int icar_thompson_init(icar_ctx* ctx, const icar_mp_options* opt) {
    if (!ctx) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    icar_mp_options d;
    icar_default_mp_options(&d);
    if (opt) d = *opt;
    static_assert(sizeof(icar_mp_options) == sizeof(ThOptions), "icar_mp_options and ThOptions must match");
    ThOptions o;
    std::memcpy(&o, &d, sizeof(o));
    if (!(o.Nt_c > 0.f) || !(o.am_s > 0.f) || !(o.rho_g > 0.f))
        return fail(ctx, ICAR_ERR_INVALID, "icar_thompson_init: Nt_c, am_s and rho_g must be positive");
    std::vector<double> bins(1000);
    th_make_constants(o, &ctx->thc, bins.data());
    if (!ctx->th_bins) CU(cudaMalloc((void**)&ctx->th_bins, 1000 * sizeof(double)));
    if (!ctx->th_aux) CU(cudaMalloc((void**)&ctx->th_aux, 800 * sizeof(double)));
    if (!ctx->th_diag) {
        CU(cudaMalloc((void**)&ctx->th_diag, sizeof(unsigned long long)));
        CU(cudaMemsetAsync(ctx->th_diag, 0, sizeof(unsigned long long), ctx->stream));
    }
    for (int t = 0; t < TH_NTABLES; ++t) {
        if (!ctx->th_tab[t]) CU(cudaMalloc((void**)&ctx->th_tab[t], th_table_elems(t) * sizeof(double)));
        CU(cudaMemsetAsync(ctx->th_tab[t], 0, th_table_elems(t) * sizeof(double), ctx->stream));   // :640-714
        ctx->tht.t[t] = ctx->th_tab[t];
    }
    CU(cudaMemcpyAsync(ctx->th_bins, bins.data(), 1000 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->launches += th_build_tables(ctx->thc, ctx->th_bins, ctx->th_aux, ctx->th_tab, ctx->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->th_ready = true;
    drop_step_graph(ctx);      // the constants are kernel parameters of a captured step
    return ICAR_OK;
}

extern "C" int icar_thompson_table(icar_ctx* ctx, int id, double* host, long long* count) {
    if (!ctx || id < 0 || id >= TH_NTABLES) return fail(ctx, ICAR_ERR_INVALID, "icar_thompson_table: bad table id");
    if (count) *count = (long long)th_table_elems(id);
    if (!host) return ICAR_OK;
    if (!ctx->th_ready) return fail(ctx, ICAR_ERR_INVALID, "icar_thompson_table before icar_thompson_init");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(host, ctx->th_tab[id], th_table_elems(id) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ICAR_OK;
}

extern "C" int icar_mp_gt_driver(icar_ctx* ctx, float* qv, float* qc, float* qr, float* qi, float* qs, float* qg,
                                 float* ni, float* nr, float* th, const float* pii, const float* p, const float* dz,
                                 float dt_in, int itimestep, float* RAINNC, const float* RAINNCV, float* SNOWNC,
                                 float* SNOWNCV, float* GRAUPELNC, float* GRAUPELNCV, float* SR, int ids, int ide,
                                 int jds, int jde, int kds, int kde, int ims, int ime, int jms, int jme, int kms,
                                 int kme, int its, int ite, int jts, int jte, int kts, int kte) {
    (void)itimestep; (void)ids; (void)jds; (void)kds; (void)kde;      // unused by the reference too
    if (!ctx || !qv || !qc || !qr || !qi || !qs || !qg || !ni || !nr || !th || !pii || !p || !dz || !RAINNC || !SR)
        return fail(ctx, ICAR_ERR_INVALID, "icar_mp_gt_driver: null array");
    if (!check_dims(ctx, ims, ime, kms, kme, jms, jme)) return fail(ctx, ICAR_ERR_INVALID, "icar_mp_gt_driver: bounds do not match the context's tile");
    if (its < ims || ite > ime || jts < jms || jte > jme || kts < kms || kte > kme)
        return fail(ctx, ICAR_ERR_INVALID, "icar_mp_gt_driver: tile outside memory bounds");
    if (!ctx->th_ready) return fail(ctx, ICAR_ERR_INVALID, "icar_mp_gt_driver before icar_thompson_init");
    CU(cudaSetDevice(ctx->device));
    const Geom& g = ctx->g;
    void* b[19];
    for (int i = 0; i < 19; ++i) { int rc = get_scratch(ctx, 80 + i, sizeof(float), &b[i]); if (rc) return rc; }
    const size_t rows = (size_t)g.nz * g.ny;
    const float* in3[12] = {qv, qc, qr, qi, qs, qg, ni, nr, th, pii, p, dz};
    int rc;
    for (int i = 0; i < 12; ++i)
        if ((rc = copy_in(ctx, b[i], in3[i], g.nx, g.ld, rows, sizeof(float)))) return rc;
    const float* in2[7] = {RAINNC, RAINNCV, SNOWNC, SNOWNCV, GRAUPELNC, GRAUPELNCV, SR};
    for (int i = 0; i < 7; ++i)
        if (in2[i] && (rc = copy_in(ctx, b[12 + i], in2[i], g.nx, ctx->ld2, g.ny, sizeof(float)))) return rc;
    MpTiles tiles;
    tiles.n = 1;
    // i_end = MIN(ite, ide-1), j_end = MIN(jte, jde-1)  (:821-822)
    tiles.its[0] = its - ims; tiles.ite[0] = std::min(ite, ide - 1) - ims;
    tiles.jts[0] = jts - jms; tiles.jte[0] = std::min(jte, jde - 1) - jms;
    tiles.along_y[0] = 0;
    auto d2 = [&](int i) { return in2[i] ? (float*)b[12 + i] : nullptr; };
    int nl = th_driver_launch(g, ctx->thc, ctx->tht, (float*)b[0], (float*)b[1], (float*)b[2], (float*)b[3], (float*)b[4],
                              (float*)b[5], (float*)b[6], (float*)b[7], (float*)b[8], (float*)b[9], (float*)b[10],
                              (float*)b[11], dt_in, d2(0), d2(1), d2(2), d2(3), d2(4), d2(5), d2(6), ctx->ld2, tiles,
                              kts - kms, kte - kms, ctx->th_diag, ctx->stream);
    if (nl < 0) return fail(ctx, ICAR_ERR_INVALID, "mp_thompson: more than 64 levels are not built");
    ctx->launches += nl;
    CU(cudaGetLastError());
    float* out3[9] = {qv, qc, qr, qi, qs, qg, ni, nr, th};
    for (int i = 0; i < 9; ++i)
        if ((rc = copy_out(ctx, out3[i], b[i], g.nx, g.ld, rows, sizeof(float)))) return rc;
    float* out2[7] = {RAINNC, nullptr, SNOWNC, SNOWNCV, GRAUPELNC, GRAUPELNCV, SR};
    for (int i = 0; i < 7; ++i)
        if (out2[i] && (rc = copy_out(ctx, out2[i], b[12 + i], g.nx, ctx->ld2, g.ny, sizeof(float)))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return ICAR_OK;
}

// ============================================================================ per-kernel profile

extern "C" int icar_profile_enable(icar_ctx* ctx, int on) {
    if (!ctx) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->aux));
    for (auto& r : ctx->prof_pending) { if (r.kid >= 0) ctx->prof_pool.push_back(r.a); ctx->prof_pool.push_back(r.b); }
    ctx->prof_pending.clear();
    for (int k = 0; k < 16; ++k) { ctx->prof_ms[k] = 0.0; ctx->prof_n[k] = 0; }
    ctx->prof_on = on != 0;
    return ICAR_OK;
}

extern "C" int icar_profile_read(icar_ctx* ctx, int max_entries, char (*names)[32], float* ms, long long* launches,
                                 int* n_out) {
    if (!ctx || !names || !ms || !launches || !n_out) return fail(ctx, ICAR_ERR_INVALID, "icar_profile_read: null argument");
    CU(cudaSetDevice(ctx->device));
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->aux));
    for (auto& r : ctx->prof_pending) {
        const int kid = r.kid >= 0 ? r.kid : -(r.kid + 100);
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ctx->prof_ms[kid] += t; ctx->prof_n[kid] += 1; }
        if (r.kid >= 0) ctx->prof_pool.push_back(r.a);      // shared start events belong to the previous record
        ctx->prof_pool.push_back(r.b);
    }
    ctx->prof_pending.clear();
    int n = 0;
    for (int k = 0; k < K_COUNT && n < max_entries; ++k) {
        if (!ctx->prof_n[k]) continue;
        std::snprintf(names[n], 32, "%s", kProfNames[k]);
        ms[n] = (float)ctx->prof_ms[k];
        launches[n] = ctx->prof_n[k];
        ++n;
    }
    *n_out = n;
    for (int k = 0; k < 16; ++k) { ctx->prof_ms[k] = 0.0; ctx->prof_n[k] = 0; }
    return ICAR_OK;
}

// ============================================================================ rest of the step loop

extern "C" int icar_diagnostic_update(icar_ctx* ctx) {
    if (!ctx) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = need_uploaded(ctx, {ICAR_VAR_PRESSURE, ICAR_VAR_POTENTIAL_TEMPERATURE}, "icar_diagnostic_update");
    if (rc) return rc;
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    ctx->launches += diagnostic_update_launch(ctx->g, (const float*)ctx->vars[ICAR_VAR_PRESSURE],
                                              (const float*)ctx->vars[ICAR_VAR_POTENTIAL_TEMPERATURE],
                                              (float*)ctx->vars[ICAR_VAR_EXNER], (float*)ctx->vars[ICAR_VAR_DENSITY],
                                              ctx->stream);
    CU(cudaGetLastError());
    ctx->uploaded[ICAR_VAR_EXNER] = true;
    ctx->uploaded[ICAR_VAR_DENSITY] = true;
    return ICAR_OK;
}

extern "C" int icar_diagnostic_output(icar_ctx* ctx, const float* dzdx, const float* dzdy, float* temperature,
                                      float* pressure_interface, float* temperature_interface, float* surface_pressure,
                                      float* u_mass, float* v_mass, float* w_real) {
    if (!ctx) return ICAR_ERR_INVALID;
    if (ctx->exchange_in_flight)
        return fail(ctx, ICAR_ERR_INVALID, "icar_diagnostic_output between icar_halo_send and icar_halo_retrieve");
    if (w_real && (!dzdx || !dzdy)) return fail(ctx, ICAR_ERR_INVALID, "icar_diagnostic_output: w_real needs dzdx and dzdy");
    CU(cudaSetDevice(ctx->device));
    const Geom& g = ctx->g;
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    // results are staged in three cell-shaped scratch fields at a time and copied out in the reference's dense layout
    constexpr size_t kSlot = 24;                   // above the slots of the leaf calls
    float* st[3] = {nullptr, nullptr, nullptr};
    for (int n = 0; n < 3; ++n) {
        void* p = nullptr;
        int rc = get_scratch(ctx, kSlot + n, sizeof(float), &p);
        if (rc) return rc;
        st[n] = (float*)p;
    }
    const size_t rows = (size_t)g.nz * g.ny;
    if (temperature || pressure_interface || temperature_interface || surface_pressure) {
        int rc = need_uploaded(ctx, {ICAR_VAR_PRESSURE, ICAR_VAR_POTENTIAL_TEMPERATURE, ICAR_VAR_EXNER}, "icar_diagnostic_output");
        if (rc) return rc;
        float* pi_dev = (pressure_interface || surface_pressure) ? st[1] : nullptr;
        ctx->launches += diagnostic_thermo_launch(g, (const float*)ctx->vars[ICAR_VAR_PRESSURE],
                                                  (const float*)ctx->vars[ICAR_VAR_POTENTIAL_TEMPERATURE],
                                                  (const float*)ctx->vars[ICAR_VAR_EXNER], temperature ? st[0] : nullptr,
                                                  pi_dev, temperature_interface ? st[2] : nullptr, ctx->stream);
        CU(cudaGetLastError());
        if (temperature && (rc = copy_out(ctx, temperature, st[0], g.nx, g.ld, rows, sizeof(float)))) return rc;
        if (pressure_interface && (rc = copy_out(ctx, pressure_interface, st[1], g.nx, g.ld, rows, sizeof(float)))) return rc;
        if (temperature_interface && (rc = copy_out(ctx, temperature_interface, st[2], g.nx, g.ld, rows, sizeof(float)))) return rc;
        if (surface_pressure)           // psfc = pressure_i(:, kms, :): one x-line per y index (time_step.f90:93)
            CU(cudaMemcpy2DAsync(surface_pressure, (size_t)g.nx * sizeof(float), st[1], g.ps() * sizeof(float),
                                 (size_t)g.nx * sizeof(float), (size_t)g.ny, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (u_mass || v_mass || w_real) {
        int rc = need_uploaded(ctx, {ICAR_VAR_U, ICAR_VAR_V}, "icar_diagnostic_output");
        if (rc) return rc;
        float *sx = nullptr, *sy = nullptr;
        if (w_real) {
            if ((rc = need_uploaded(ctx, {ICAR_VAR_W, ICAR_VAR_JACOBIAN}, "icar_diagnostic_output"))) return rc;
            void* p = nullptr;
            if ((rc = get_scratch(ctx, kSlot + 3, sizeof(float), &p))) return rc;
            sx = (float*)p;
            if ((rc = get_scratch(ctx, kSlot + 4, sizeof(float), &p))) return rc;
            sy = (float*)p;
            if ((rc = copy_in(ctx, sx, dzdx, g.nx + 1, g.ldu, rows, sizeof(float)))) return rc;
            if ((rc = copy_in(ctx, sy, dzdy, g.nx, g.ld, (size_t)g.nz * (g.ny + 1), sizeof(float)))) return rc;
        }
        ctx->launches += diagnostic_winds_launch(g, (const float*)ctx->vars[ICAR_VAR_U], (const float*)ctx->vars[ICAR_VAR_V],
                                                 (const float*)ctx->vars[ICAR_VAR_W], sx, sy,
                                                 (const float*)ctx->vars[ICAR_VAR_JACOBIAN], u_mass ? st[0] : nullptr,
                                                 v_mass ? st[1] : nullptr, w_real ? st[2] : nullptr, ctx->stream);
        CU(cudaGetLastError());
        if (u_mass && (rc = copy_out(ctx, u_mass, st[0], g.nx, g.ld, rows, sizeof(float)))) return rc;
        if (v_mass && (rc = copy_out(ctx, v_mass, st[1], g.nx, g.ld, rows, sizeof(float)))) return rc;
        if (w_real && (rc = copy_out(ctx, w_real, st[2], g.nx, g.ld, rows, sizeof(float)))) return rc;
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return ICAR_OK;
}

extern "C" int icar_compute_dt(icar_ctx* ctx, const float* dz_levels, float cfl, int strictness, double* dt_seconds) {
    if (!ctx || !dz_levels || !dt_seconds) return fail(ctx, ICAR_ERR_INVALID, "icar_compute_dt: null argument");
    if (strictness < 1 || strictness > 5) return fail(ctx, ICAR_ERR_INVALID, "icar_compute_dt: cfl_strictness must be 1..5");
    CU(cudaSetDevice(ctx->device));
    int rc = need_uploaded(ctx, {ICAR_VAR_U, ICAR_VAR_V, ICAR_VAR_W}, "icar_compute_dt");
    if (rc) return rc;
    const Geom& g = ctx->g;
    if (!ctx->cfl_max) {
        CU(cudaMalloc((void**)&ctx->cfl_max, 4 * sizeof(unsigned)));
        CU(cudaMalloc((void**)&ctx->cfl_dt, sizeof(double)));
        CU(cudaMalloc((void**)&ctx->cfl_dz, (size_t)g.nz * sizeof(float)));
    }
    CU(cudaMemcpyAsync(ctx->cfl_dz, dz_levels, (size_t)g.nz * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->cfl_max, 0, 4 * sizeof(unsigned), ctx->stream));
    ctx->launches += cfl_maxima_launch(g, (const float*)ctx->vars[ICAR_VAR_U], (const float*)ctx->vars[ICAR_VAR_V],
                                       (const float*)ctx->vars[ICAR_VAR_W], ctx->cfl_dz, ctx->opt.dx, ctx->cfl_max,
                                       ctx->stream);
    CU(cudaGetLastError());
    unsigned bits[4];
    CU(cudaMemcpyAsync(bits, ctx->cfl_max, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float m[4];
    std::memcpy(m, bits, sizeof(m));
    // time_step.f90:234-318 (REAL arithmetic)
    const float sqrt3 = std::sqrt(3.0f) * 1.001f, three_d_cfl = 0.577350269f;
    float maxwind3d = 0.f, maxwind1d = 0.f;
    if (strictness == 1) {
        maxwind1d = std::max(std::max(m[1], m[2]), m[3]);
        maxwind3d = maxwind1d * sqrt3;
    } else if (strictness == 5) {
        maxwind3d = m[1] + m[2] + m[3];
    } else {
        maxwind3d = m[0];
        if (strictness == 2) {
            maxwind3d = maxwind3d * three_d_cfl;
            maxwind1d = std::max(std::max(m[1], m[2]), m[3]);
            maxwind3d = std::max(maxwind1d, maxwind3d);
        } else if (strictness == 4) {
            maxwind3d = maxwind3d * sqrt3;
        }
    }
    const float dt = cfl / maxwind3d;
    double seconds = (double)dt;
    if (ctx->nimages > 1 && ctx->comm) {                        // call co_min(seconds), :413
        // (the collective comes BEFORE the "too small" stop: an image that returned early would leave the others
        // blocked in it; this way every image sees the same minimum and stops, or goes on, together)
        CU(cudaMemcpyAsync(ctx->cfl_dt, &seconds, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        NC(g_nccl.AllReduce(ctx->cfl_dt, ctx->cfl_dt, 1, ncclDouble, ncclMin, ctx->comm, ctx->stream));
        CU(cudaMemcpyAsync(&seconds, ctx->cfl_dt, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    } else if (ctx->nimages > 1 && !ctx->linked) {
        return fail(ctx, ICAR_ERR_COMM, "icar_compute_dt: icar_comm_init was not called");
    }
    if (seconds < 1e-1) {                                       // :322-328
        char msg[240];
        std::snprintf(msg, sizeof msg, "ERROR time step too small: dt = %g (this image: dt = %g, Umax = %g, Vmax = %g, Wmax = %g)",
                      seconds, dt, m[1], m[2], m[3]);
        return fail(ctx, ICAR_ERR_PHYSICS, msg);
    }
    *dt_seconds = std::min(seconds, 120.0);                     // :421
    return ICAR_OK;
}

extern "C" int icar_enforce_limits(icar_ctx* ctx) {
    if (!ctx) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    float* fields[ICAR_MAX_EXCHANGED];
    for (int s = 0; s < ctx->nexch; ++s) fields[s] = (float*)ctx->vars[ctx->exch[s]];
    ctx->launches += enforce_limits_launch(ctx->g, fields, ctx->nexch, ctx->stream);
    CU(cudaGetLastError());
    return ICAR_OK;
}

extern "C" int icar_upload_tendency(icar_ctx* ctx, int var, const float* host) {
    if (!ctx || !host || var < 0 || var >= ICAR_VAR_COUNT || var_shape(var) == D2)
        return fail(ctx, ICAR_ERR_INVALID, "icar_upload_tendency: bad argument (3-D REAL variables only)");
    CU(cudaSetDevice(ctx->device));
    int width, pitch; size_t rows;
    var_geometry(ctx, var, &width, &pitch, &rows);
    if (!ctx->dqdt[var]) {
        const size_t bytes = (var_shape(var) == USTAG) ? (size_t)ctx->g.ldu * ctx->g.nz * ctx->g.ny * sizeof(float)
                                                       : cell_elems(ctx) * sizeof(float);
        CU(cudaMalloc((void**)&ctx->dqdt[var], bytes));
        CU(cudaMemsetAsync(ctx->dqdt[var], 0, bytes, ctx->stream));
    }
    int rc = copy_in(ctx, ctx->dqdt[var], host, width, pitch, rows, sizeof(float));
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return ICAR_OK;
}

extern "C" int icar_apply_forcing(icar_ctx* ctx, double dt_seconds) {
    if (!ctx) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    const icar_grid& gr = ctx->grid;
    // this%west_boundary = (ximg == 1) ... (domain_obj.f90:2211-2214): the tile touches that domain edge
    const bool edges[4] = {gr.ims == gr.ids, gr.ime == gr.ide, gr.jms == gr.jds, gr.jme == gr.jde};
    for (int var = 0; var < ICAR_VAR_COUNT; ++var) {
        if (!ctx->dqdt[var]) continue;
        const bool diagnosed = (var == ICAR_VAR_U || var == ICAR_VAR_V || var == ICAR_VAR_W || var == ICAR_VAR_PRESSURE);
        int width, pitch; size_t rows;
        var_geometry(ctx, var, &width, &pitch, &rows);
        ctx->launches += apply_forcing_launch((float*)ctx->vars[var], ctx->dqdt[var], width, ctx->g.nz,
                                              (int)(rows / ctx->g.nz), pitch, dt_seconds, !diagnosed, edges, ctx->stream);
    }
    CU(cudaGetLastError());
    return ICAR_OK;
}

// step(domain, end_time, options), src/main/time_step.f90:438-549: the loop over internal time steps up to `end_time`
// with everything on the device.  Model time is REAL*8 seconds here (the reference's Time_type holds a modified Julian
// day and adds dt%days(); the two round differently below a microsecond, which is what the dt > 1e-3 test of :484 is
// there to absorb).
extern "C" int icar_step(icar_ctx* ctx, double end_time, const float* dz_levels, float cfl_reduction_factor,
                         int cfl_strictness, int* steps_taken) {
    if (!ctx) return ICAR_ERR_INVALID;
    if (steps_taken) *steps_taken = 0;
    int rc;
    while (ctx->model_time < end_time) {                                                  // :464
        double dt = 0.0;
        if ((rc = icar_compute_dt(ctx, dz_levels, cfl_reduction_factor, cfl_strictness, &dt))) return rc;   // update_dt, :466
        if (ctx->model_time + dt > end_time) dt = end_time - ctx->model_time;             // :469-471
        if ((rc = icar_diagnostic_update(ctx))) return rc;                                // :474
        const double t0 = ctx->model_time;
        if (dt > 1e-3) {                                                                  // :484
            // rad / lsm / pbl / convect (:489-509) are 0 in every configuration of this path
            if ((rc = icar_physics_step(ctx, (float)dt))) return rc;                      // :512-529, real(dt%seconds())
            ctx->model_time = t0;          // icar_physics_step advanced it by the REAL*4 dt; the loop keeps REAL*8 (:547)
            if ((rc = icar_apply_forcing(ctx, dt))) return rc;                            // :537
            if ((end_time - t0) < dt * 2.0)                                               // :540-542
                if ((rc = icar_enforce_limits(ctx))) return rc;
        }
        ctx->model_time = t0 + dt;                                                        // :547
        if (steps_taken) ++*steps_taken;
    }
    return ICAR_OK;
}

extern "C" int icar_balance_uvw(icar_ctx* ctx) {
    if (!ctx) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = need_uploaded(ctx, {ICAR_VAR_U, ICAR_VAR_V, ICAR_VAR_JACOBIAN_U, ICAR_VAR_JACOBIAN_V, ICAR_VAR_JACOBIAN_W,
                                 ICAR_VAR_ADVECTION_DZ}, "icar_balance_uvw");
    if (rc) return rc;
    ctx->launches += balance_uvw_launch(ctx->g, (const float*)ctx->vars[ICAR_VAR_U], (const float*)ctx->vars[ICAR_VAR_V],
                                        (float*)ctx->vars[ICAR_VAR_W], (const float*)ctx->vars[ICAR_VAR_JACOBIAN_U],
                                        (const float*)ctx->vars[ICAR_VAR_JACOBIAN_V],
                                        (const float*)ctx->vars[ICAR_VAR_JACOBIAN_W],
                                        (const float*)ctx->vars[ICAR_VAR_ADVECTION_DZ], ctx->opt.dx, ctx->stream);
    CU(cudaGetLastError());
    ctx->uploaded[ICAR_VAR_W] = true;
    return ICAR_OK;
}

// ------------------------------------------------------------------ linear-theory winds ("next" rank 3)

extern "C" void icar_add_buffer_topo(const float* terrain, int nx, int ny, int smooth_window, int buffer, double* out) {
    if (terrain && out && nx > 0 && ny > 0 && buffer >= 0) lw_add_buffer_topo(terrain, nx, ny, smooth_window, buffer, out);
}

extern "C" int icar_linear_winds_init(icar_ctx* ctx, const float* global_terrain, int nx_global, int ny_global,
                                      int buffer) {
    if (!ctx) return ICAR_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    const icar_grid& gr = ctx->grid;
    if (nx_global != gr.ide - gr.ids + 1 || ny_global != gr.jde - gr.jds + 1)
        return fail(ctx, ICAR_ERR_INVALID, "icar_linear_winds_init: global_terrain must cover ids:ide x jds:jde");
    lw_destroy(ctx->lw);
    ctx->lw = nullptr;
    std::string err;
    const int rc = lw_create(&ctx->lw, global_terrain, nx_global, ny_global, buffer, ctx->opt.dx, ctx->stream, &err);
    if (rc) return fail(ctx, rc == 1 ? ICAR_ERR_INVALID : ICAR_ERR_CUDA, "icar_linear_winds_init: " + err);
    ctx->launches += 1;
    return ICAR_OK;
}

extern "C" int icar_linear_winds_dims(icar_ctx* ctx, int* fft_nx, int* fft_ny, int* buffer) {
    if (!ctx || !ctx->lw) return fail(ctx, ICAR_ERR_INVALID, "icar_linear_winds_dims: call icar_linear_winds_init first");
    lw_dims(ctx->lw, fft_nx, fft_ny, buffer);
    return ICAR_OK;
}

extern "C" int icar_linear_winds_terrain_frequency(icar_ctx* ctx, double* host) {
    if (!ctx || !ctx->lw || !host)
        return fail(ctx, ICAR_ERR_INVALID, "icar_linear_winds_terrain_frequency: call icar_linear_winds_init first");
    CU(cudaSetDevice(ctx->device));
    std::string err;
    if (lw_terrain_frequency(ctx->lw, host, &err)) return fail(ctx, ICAR_ERR_CUDA, err);
    return ICAR_OK;
}

extern "C" int icar_linear_perturbation(icar_ctx* ctx, float U, float V, float Nsq, float z_bottom, float z_top,
                                        const float* z_bottom_2d, const float* z_top_2d, float minimum_step,
                                        double* u_perturb, double* v_perturb) {
    if (!ctx || !ctx->lw) return fail(ctx, ICAR_ERR_INVALID, "icar_linear_perturbation: call icar_linear_winds_init first");
    if ((z_bottom_2d == nullptr) != (z_top_2d == nullptr))
        return fail(ctx, ICAR_ERR_INVALID, "icar_linear_perturbation: z_bottom_2d and z_top_2d go together");
    CU(cudaSetDevice(ctx->device));
    std::string err;
    const int rc = z_bottom_2d ? lw_perturbation_varyingz(ctx->lw, U, V, Nsq, z_bottom_2d, z_top_2d, minimum_step,
                                                          &ctx->launches, &err)
                               : lw_perturbation_constz(ctx->lw, U, V, Nsq, z_bottom, z_top, minimum_step,
                                                        &ctx->launches, &err);
    if (rc) return fail(ctx, rc == 1 ? ICAR_ERR_INVALID : rc == 3 ? ICAR_ERR_PHYSICS : ICAR_ERR_CUDA, "icar_linear_perturbation: " + err);
    if (u_perturb || v_perturb)
        if (lw_download_perturbation(ctx->lw, u_perturb, v_perturb, &err)) return fail(ctx, ICAR_ERR_CUDA, err);
    return ICAR_OK;
}

extern "C" int icar_linear_perturb(icar_ctx* ctx, float U, float V, float Nsq, const float* global_z_interface,
                                   const float* global_dz_interface, const float* layer_height, const float* dz_levels,
                                   float minimum_layer_size, float linear_contribution, float linear_update_fraction) {
    if (!ctx || !ctx->lw) return fail(ctx, ICAR_ERR_INVALID, "icar_linear_perturb: call icar_linear_winds_init first");
    const bool varying = global_z_interface != nullptr;
    if (varying ? !global_dz_interface : !(layer_height && dz_levels))
        return fail(ctx, ICAR_ERR_INVALID, "icar_linear_perturb: give global_z_interface + global_dz_interface "
                                           "(space_varying_dz) or layer_height + dz_levels");
    CU(cudaSetDevice(ctx->device));
    int rc = need_uploaded(ctx, {ICAR_VAR_U, ICAR_VAR_V}, "icar_linear_perturb");
    if (rc) return rc;
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    const icar_grid& gr = ctx->grid;
    const Geom& g = ctx->g;
    int gnx = 0, gny = 0;
    const float* terrain = lw_terrain(ctx->lw, &gnx, &gny);
    std::vector<float> zb, zt;
    if (varying) { zb.resize((size_t)gnx * gny); zt.resize((size_t)gnx * gny); }
    std::string err;
    for (int k = 0; k < g.nz; ++k) {
        if (varying) {
            // z_interface(:, z, :) - terrain and ... + dz_interface(:, z, :)   (linear_winds.f90:723-726)
            for (int j = 0; j < gny; ++j)
                for (int i = 0; i < gnx; ++i) {
                    const size_t c3 = (size_t)i + (size_t)gnx * ((size_t)k + (size_t)g.nz * (size_t)j);
                    const float b = global_z_interface[c3] - terrain[(size_t)j * gnx + i];
                    zb[(size_t)j * gnx + i] = b;
                    zt[(size_t)j * gnx + i] = b + global_dz_interface[c3];
                }
            rc = lw_perturbation_varyingz(ctx->lw, U, V, Nsq, zb.data(), zt.data(), minimum_layer_size, &ctx->launches, &err);
        } else {
            const float half = dz_levels[k] / 2.0f;                  // :730-732
            rc = lw_perturbation_constz(ctx->lw, U, V, Nsq, layer_height[k] - half, layer_height[k] + half,
                                        minimum_layer_size, &ctx->launches, &err);
        }
        if (!rc)
            rc = lw_apply_level(ctx->lw, g, gr.ims - gr.ids, gr.jms - gr.jds, k, (float*)ctx->vars[ICAR_VAR_U],
                                (float*)ctx->vars[ICAR_VAR_V], linear_contribution, linear_update_fraction,
                                &ctx->launches, &err);
        if (rc) return fail(ctx, rc == 1 ? ICAR_ERR_INVALID : rc == 3 ? ICAR_ERR_PHYSICS : ICAR_ERR_CUDA, "icar_linear_perturb: " + err);
    }
    return ICAR_OK;
}

extern "C" void icar_default_lt_options(icar_lt_options* o) {
    if (!o) return;
    // options_obj.f90:1448-1479
    o->variable_N = 1; o->smooth_nsq = 1; o->vert_smooth = 10; o->stability_window_size = 10;
    o->max_stability = 6e-4f; o->min_stability = 1e-7f; o->N_squared = 3e-5f;
    o->linear_contribution = 1.f; o->linear_update_fraction = 0.2f;
    o->dirmax = 2.f * 3.1415927f; o->dirmin = 0.f; o->spdmax = 30.f; o->spdmin = 0.f;
    o->n_dir_values = 24; o->n_nsq_values = 5; o->n_spd_values = 6;
    o->minimum_layer_size = 100.f;
}

extern "C" int icar_linear_winds_set_z(icar_ctx* ctx, const float* z) {
    if (!ctx || !ctx->lw) return fail(ctx, ICAR_ERR_INVALID, "icar_linear_winds_set_z: call icar_linear_winds_init first");
    CU(cudaSetDevice(ctx->device));
    std::string err;
    const int rc = lw_set_z(ctx->lw, ctx->g, z, &err);
    if (rc) return fail(ctx, rc == 1 ? ICAR_ERR_INVALID : ICAR_ERR_CUDA, "icar_linear_winds_set_z: " + err);
    return ICAR_OK;
}

extern "C" int icar_linear_perturb_spatial(icar_ctx* ctx, const icar_lt_options* opt, const float* global_z_interface,
                                           const float* global_dz_interface, const float* layer_height,
                                           const float* dz_levels, long long* entries_computed) {
    if (!ctx || !ctx->lw) return fail(ctx, ICAR_ERR_INVALID, "icar_linear_perturb_spatial: call icar_linear_winds_init first");
    CU(cudaSetDevice(ctx->device));
    int rc = need_uploaded(ctx, {ICAR_VAR_U, ICAR_VAR_V, ICAR_VAR_POTENTIAL_TEMPERATURE, ICAR_VAR_EXNER, ICAR_VAR_WATER_VAPOR},
                           "icar_linear_perturb_spatial");
    if (rc) return rc;
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    icar_lt_options d;
    icar_default_lt_options(&d);
    if (opt) d = *opt;
    static_assert(sizeof(icar_lt_options) == sizeof(LtOptions), "icar_lt_options and LtOptions must match");
    LtOptions o;
    std::memcpy(&o, &d, sizeof(o));
    auto have = [&](int v) -> const float* { return ctx->uploaded[v] ? (const float*)ctx->vars[v] : nullptr; };
    // hydrometeors = cloud water + cloud ice + rain + snow, whichever are associated (:905-913)
    LtFields f{(const float*)ctx->vars[ICAR_VAR_POTENTIAL_TEMPERATURE], (const float*)ctx->vars[ICAR_VAR_EXNER],
               (const float*)ctx->vars[ICAR_VAR_WATER_VAPOR], have(ICAR_VAR_CLOUD_WATER),
               ctx->opt.microphysics == 1 ? have(ICAR_VAR_CLOUD_ICE) : nullptr, have(ICAR_VAR_RAIN_IN_AIR),
               have(ICAR_VAR_SNOW_IN_AIR)};
    const icar_grid& gr = ctx->grid;
    std::string err;
    rc = lw_spatial_winds(ctx->lw, ctx->g, gr.ims - gr.ids, gr.jms - gr.jds, f, o, (float*)ctx->vars[ICAR_VAR_U],
                          (float*)ctx->vars[ICAR_VAR_V], global_z_interface, global_dz_interface, layer_height, dz_levels,
                          entries_computed, &ctx->launches, &err);
    if (rc) return fail(ctx, rc == 1 ? ICAR_ERR_INVALID : rc == 3 ? ICAR_ERR_PHYSICS : ICAR_ERR_CUDA, "icar_linear_perturb_spatial: " + err);
    return ICAR_OK;
}

// ------------------------------------------------------------------ restart / output files ("next" rank 4)
//
// One self-describing little-endian file per image (the reference writes one NetCDF file per image as well,
// output_obj.f90; there is no NetCDF library in this image).  Layout:
//   char magic[8] = "ICARB2RS"; int32 version = 1; int32 nvars; int32 grid[19] (icar_grid member order);
//   int32 microphysics; double model_time; double last_model_time (mp()'s update-interval memory);
//   nvars x { int32 var (enum icar_var); int32 elem_size; int32 dims[3] = (nx', ny', nz') in FILE order, nz' = 1
//             for 2-D variables; int64 nbytes; data }
// 3-D data are in the reference's file order (x, y, z) = reshape(data_3d, order=[1,3,2]) (output_obj.f90:423-426).

namespace {

const char kRestartMagic[8] = {'I', 'C', 'A', 'R', 'B', '2', 'R', 'S'};

struct RestartHeader {
    char magic[8];
    int version, nvars;
    int grid[19];
    int microphysics;
    double model_time, last_model_time;
};
struct RestartVar {
    int var, elem_size, dims[3];
    long long nbytes;
};

bool read_exact(FILE* f, void* p, size_t n) { return std::fread(p, 1, n, f) == n; }

int read_restart_header(FILE* f, RestartHeader* h, std::string* err) {
    if (!read_exact(f, h->magic, 8) || std::memcmp(h->magic, kRestartMagic, 8) != 0) { *err = "not an icar_b200 restart file"; return 1; }
    if (!read_exact(f, &h->version, 4) || !read_exact(f, &h->nvars, 4) || !read_exact(f, h->grid, sizeof h->grid) ||
        !read_exact(f, &h->microphysics, 4) || !read_exact(f, &h->model_time, 8) || !read_exact(f, &h->last_model_time, 8)) {
        *err = "truncated restart header";
        return 1;
    }
    if (h->version != 1) { *err = "restart file version " + std::to_string(h->version) + " (this library reads 1)"; return 1; }
    if (h->nvars < 0 || h->nvars > ICAR_VAR_COUNT) { *err = "restart file: bad variable count"; return 1; }
    return 0;
}

int read_restart_var(FILE* f, RestartVar* v, std::string* err) {
    if (!read_exact(f, &v->var, 4) || !read_exact(f, &v->elem_size, 4) || !read_exact(f, v->dims, 12) ||
        !read_exact(f, &v->nbytes, 8)) { *err = "truncated restart file (variable record)"; return 1; }
    if (v->var < 0 || v->var >= ICAR_VAR_COUNT || (v->elem_size != 4 && v->elem_size != 8) ||
        v->nbytes != (long long)v->elem_size * v->dims[0] * v->dims[1] * v->dims[2]) { *err = "corrupt restart file (variable record)"; return 1; }
    return 0;
}

}  // namespace

extern "C" int icar_restart_info(const char* path, icar_grid* grid, int* microphysics, double* model_time, int* nvars,
                                 int* vars) {
    if (!path) return ICAR_ERR_INVALID;
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(nullptr, ICAR_ERR_INVALID, std::string("icar_restart_info: cannot open ") + path);
    RestartHeader h;
    std::string err;
    int rc = read_restart_header(f, &h, &err);
    for (int n = 0; !rc && n < h.nvars; ++n) {
        RestartVar v;
        rc = read_restart_var(f, &v, &err);
        if (!rc && std::fseek(f, (long)v.nbytes, SEEK_CUR) != 0) { err = "truncated restart file (data)"; rc = 1; }
        if (!rc && vars) vars[n] = v.var;
    }
    if (!rc) {
        // the data of the last variable must be there too
        const long here = std::ftell(f);
        std::fseek(f, 0, SEEK_END);
        if (std::ftell(f) < here) { err = "truncated restart file (data)"; rc = 1; }
    }
    std::fclose(f);
    if (rc) return fail(nullptr, ICAR_ERR_INVALID, std::string("icar_restart_info: ") + err);
    if (grid) std::memcpy(grid, h.grid, sizeof h.grid);
    if (microphysics) *microphysics = h.microphysics;
    if (model_time) *model_time = h.model_time;
    if (nvars) *nvars = h.nvars;
    return ICAR_OK;
}

extern "C" int icar_write_restart(icar_ctx* ctx, const char* path) {
    if (!ctx || !path) return fail(ctx, ICAR_ERR_INVALID, "icar_write_restart: bad argument");
    static_assert(sizeof(icar_grid) == 19 * sizeof(int), "icar_grid is 19 ints");
    CU(cudaSetDevice(ctx->device));
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(ctx, ICAR_ERR_INVALID, std::string("icar_write_restart: cannot create ") + path);
    // the precipitation accumulators start at zero and are only ever written by the mp kernels: always saved
    auto saved = [&](int v) { return ctx->vars[v] && (ctx->uploaded[v] || var_shape(v) == D2); };
    int nvars = 0;
    for (int v = 0; v < ICAR_VAR_COUNT; ++v) nvars += saved(v) ? 1 : 0;
    const int version = 1;
    bool ok = std::fwrite(kRestartMagic, 1, 8, f) == 8 && std::fwrite(&version, 4, 1, f) == 1 &&
              std::fwrite(&nvars, 4, 1, f) == 1 && std::fwrite(&ctx->grid, sizeof(icar_grid), 1, f) == 1 &&
              std::fwrite(&ctx->opt.microphysics, 4, 1, f) == 1 && std::fwrite(&ctx->model_time, 8, 1, f) == 1 &&
              std::fwrite(&ctx->last_model_time, 8, 1, f) == 1;
    std::vector<char> host;
    int rc = ICAR_OK;
    for (int var = 0; ok && rc == ICAR_OK && var < ICAR_VAR_COUNT; ++var) {
        if (!saved(var)) continue;
        int width, pitch; size_t rows;
        var_geometry(ctx, var, &width, &pitch, &rows);
        const bool two_d = var_shape(var) == D2;
        const int nz = two_d ? 1 : ctx->g.nz, rows_y = (int)(rows / nz);
        RestartVar rv{var, (int)elem_size(var), {width, rows_y, nz}, (long long)elem_size(var) * width * (long long)rows};
        host.resize((size_t)rv.nbytes);
        if (two_d) {
            rc = copy_out(ctx, host.data(), ctx->vars[var], width, pitch, rows, elem_size(var));
        } else {
            void* dense = nullptr;
            rc = get_scratch(ctx, 0, sizeof(float), &dense);
            if (rc == ICAR_OK) {
                ctx->launches += output_reorder_launch((float*)ctx->vars[var], (float*)dense, width, nz, rows_y, pitch, true, ctx->stream);
                if (cudaMemcpyAsync(host.data(), dense, (size_t)rv.nbytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
                    rc = fail(ctx, ICAR_ERR_CUDA, "icar_write_restart: device to host copy failed");
            }
        }
        if (rc == ICAR_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            rc = fail(ctx, ICAR_ERR_CUDA, std::string("icar_write_restart: ") + cudaGetErrorString(cudaGetLastError()));
        if (rc == ICAR_OK)
            ok = std::fwrite(&rv.var, 4, 1, f) == 1 && std::fwrite(&rv.elem_size, 4, 1, f) == 1 &&
                 std::fwrite(rv.dims, 4, 3, f) == 3 && std::fwrite(&rv.nbytes, 8, 1, f) == 1 &&
                 std::fwrite(host.data(), 1, (size_t)rv.nbytes, f) == (size_t)rv.nbytes;
    }
    ok = (std::fclose(f) == 0) && ok;
    if (rc != ICAR_OK) return rc;
    if (!ok) return fail(ctx, ICAR_ERR_INVALID, std::string("icar_write_restart: write to ") + path + " failed");
    return ICAR_OK;
}

extern "C" int icar_read_restart(icar_ctx* ctx, const char* path) {
    if (!ctx || !path) return fail(ctx, ICAR_ERR_INVALID, "icar_read_restart: bad argument");
    CU(cudaSetDevice(ctx->device));
    { int jr = join_aux(ctx); if (jr) return jr; jr = join_winds(ctx); if (jr) return jr; }
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(ctx, ICAR_ERR_INVALID, std::string("icar_read_restart: cannot open ") + path);
    RestartHeader h;
    std::string err;
    int bad = read_restart_header(f, &h, &err);
    // restart_domain_error (restart.f90): the file must have been written for this tile and this scheme
    if (!bad && std::memcmp(h.grid, &ctx->grid, sizeof h.grid) != 0) { err = "the file was written for a different grid / decomposition"; bad = 1; }
    if (!bad && h.microphysics != ctx->opt.microphysics) { err = "the file was written with a different microphysics scheme"; bad = 1; }
    std::vector<char> host;
    int rc = ICAR_OK;
    for (int n = 0; !bad && rc == ICAR_OK && n < h.nvars; ++n) {
        RestartVar rv;
        bad = read_restart_var(f, &rv, &err);
        if (bad) break;
        int width, pitch; size_t rows;
        var_geometry(ctx, rv.var, &width, &pitch, &rows);
        const bool two_d = var_shape(rv.var) == D2;
        const int nz = two_d ? 1 : ctx->g.nz, rows_y = (int)(rows / nz);
        if (!ctx->vars[rv.var] || rv.elem_size != (int)elem_size(rv.var) || rv.dims[0] != width || rv.dims[1] != rows_y || rv.dims[2] != nz) {
            err = "variable " + std::to_string(rv.var) + " has a different shape in the file";
            bad = 1;
            break;
        }
        host.resize((size_t)rv.nbytes);
        if (!read_exact(f, host.data(), (size_t)rv.nbytes)) { err = "truncated restart file (data)"; bad = 1; break; }
        if (two_d) {
            rc = copy_in(ctx, ctx->vars[rv.var], host.data(), width, pitch, rows, elem_size(rv.var));
        } else {
            void* dense = nullptr;
            rc = get_scratch(ctx, 0, sizeof(float), &dense);
            if (rc == ICAR_OK) {
                if (cudaMemcpyAsync(dense, host.data(), (size_t)rv.nbytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
                    rc = fail(ctx, ICAR_ERR_CUDA, "icar_read_restart: host to device copy failed");
                else
                    ctx->launches += output_reorder_launch((float*)ctx->vars[rv.var], (float*)dense, width, nz, rows_y, pitch, false, ctx->stream);
            }
        }
        if (rc == ICAR_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            rc = fail(ctx, ICAR_ERR_CUDA, std::string("icar_read_restart: ") + cudaGetErrorString(cudaGetLastError()));
        if (rc == ICAR_OK) ctx->uploaded[rv.var] = true;
    }
    std::fclose(f);
    if (rc != ICAR_OK) return rc;
    if (bad) return fail(ctx, ICAR_ERR_INVALID, std::string("icar_read_restart: ") + err);
    ctx->model_time = h.model_time;
    ctx->last_model_time = h.last_model_time;
    ctx->winds_f64_valid = false;
    return ICAR_OK;
}
