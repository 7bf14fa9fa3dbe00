// api.cu -- the C ABI of libtgs.so (include/tgs.h): argument checking, saved-buffer layouts and
// the host-side orchestration of the kernels.  No torch types; the Python host layer
// (touch-gs_b200/rasterizer.py) binds these with ctypes.
#include "tgs_common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

// ------------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_own{0}, g_cub{0};

void tgs_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int tgs_check_cuda(cudaError_t e, const char* what, const char* file, int line) {
    if (e == cudaSuccess) return 0;
    tgs_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return (int)e;
}
void tgs_count_own(int n) { g_own += (uint64_t)n; }
void tgs_count_cub(int n) { g_cub += (uint64_t)n; }

// -------------------------------------------------------------------------- stage timers
#include <mutex>
namespace {
struct ProfPair { cudaEvent_t a, b; int stage; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfPair> g_prof_pairs;      // recorded since the last read
std::vector<ProfPair> g_prof_free;       // recycled events
cudaEvent_t g_prof_open[TGS_NUM_STAGES];
}  // namespace
void tgs_prof_begin(int stage, cudaStream_t st) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfPair p;
    if (!g_prof_free.empty()) { p = g_prof_free.back(); g_prof_free.pop_back(); }
    else { if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return; }
    p.stage = stage;
    cudaEventRecord(p.a, st);
    g_prof_open[stage] = p.a;
    g_prof_pairs.push_back(p);
}
void tgs_prof_end(int stage, cudaStream_t st) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto it = g_prof_pairs.rbegin(); it != g_prof_pairs.rend(); ++it)
        if (it->stage == stage && it->a == g_prof_open[stage]) { cudaEventRecord(it->b, st); break; }
}
extern "C" int tgs_profile_enable(int32_t on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    return 0;
}
extern "C" int tgs_profile_read(float* ms, int32_t* cnt) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = 0; i < TGS_NUM_STAGES; ++i) { if (ms) ms[i] = 0.f; if (cnt) cnt[i] = 0; }
    for (auto& p : g_prof_pairs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) { if (ms) ms[p.stage] += t; if (cnt) cnt[p.stage] += 1; }
        g_prof_free.push_back(p);
    }
    g_prof_pairs.clear();
    cudaGetLastError();
    return 0;
}

// ------------------------------------------------------------------------------- layouts
namespace {
struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) { size_t o = off; off = tgs_align_up(off + bytes); return o; }
};
int ceil_log2(uint32_t v) { int b = 0; while ((1u << b) < v) ++b; return b; }
}  // namespace

extern "C" int tgs_geom_layout(int32_t N, TgsGeomLayout* o) {
    Carver c; size_t n = (size_t)(N > 0 ? N : 0);
    o->records = c.take(n * sizeof(TgsRecord));
    o->cov3D = c.take(n * 6 * sizeof(float));
    o->tiles_touched = c.take(n * sizeof(uint32_t));
    o->clamped = c.take(n);
    o->rect = c.take(n * sizeof(uint2));
    o->depth_keys = c.take(n * sizeof(uint32_t));
    o->ids = c.take(n * sizeof(uint32_t));
    o->depth_keys_sorted = c.take(n * sizeof(uint32_t));
    o->order = c.take(n * sizeof(uint32_t));
    o->span_sorted = c.take(n * sizeof(uint2));
    o->temp_bytes = tgs_depth_sort_temp_bytes(N > 0 ? N : 1);
    o->temp = c.take(o->temp_bytes);
    o->total = c.off > 0 ? c.off : TGS_ALIGN;
    return 0;
}
extern "C" int tgs_binning_layout(int64_t I, TgsBinningLayout* o) {
    Carver c; size_t n = (size_t)(I > 0 ? I : 0);
    o->vals_sorted = c.take(n * sizeof(uint32_t));
    o->slots = (n >> 8) + 2;
    o->ckpt = c.take(o->slots * TGS_CKPT_FLOATS * sizeof(float));
    o->slot_tile = c.take(o->slots * sizeof(uint32_t));
    o->ckpt_list = c.take(o->slots * sizeof(uint32_t));
    o->work_counter = c.take(2 * sizeof(uint32_t));
    o->total = c.off;
    return 0;
}
extern "C" int tgs_image_layout(int32_t W, int32_t H, TgsImageLayout* o) {
    Carver c; size_t p = (size_t)W * H;
    o->final_T = c.take(p * 4);
    o->n_contrib = c.take(p * 4);
    o->depth_raw = c.take(p * 4);
    o->color_acc = c.take(p * 12);
    const size_t T = (size_t)((W + TGS_TILE - 1) / TGS_TILE) * (size_t)((H + TGS_TILE - 1) / TGS_TILE);
    o->ranges = c.take((T > 0 ? T : 1) * sizeof(uint2));
    o->count = c.take(2 * sizeof(uint32_t));
    o->total = c.off;
    return 0;
}

GeomView tgs_geom_view(void* base, int N) {
    TgsGeomLayout l; tgs_geom_layout(N, &l);
    char* b = (char*)base; GeomView v;
    v.records = (TgsRecord*)(b + l.records); v.cov3D = (float*)(b + l.cov3D);
    v.tiles_touched = (uint32_t*)(b + l.tiles_touched);
    v.clamped = (uint8_t*)(b + l.clamped); v.rect = (uint2*)(b + l.rect);
    v.depth_keys = (uint32_t*)(b + l.depth_keys); v.ids = (uint32_t*)(b + l.ids);
    v.depth_keys_sorted = (uint32_t*)(b + l.depth_keys_sorted); v.order = (uint32_t*)(b + l.order);
    v.span_sorted = (uint2*)(b + l.span_sorted);
    v.temp = b + l.temp; v.temp_bytes = l.temp_bytes;
    return v;
}
BinView tgs_bin_view(void* base, int64_t I) {
    TgsBinningLayout l; tgs_binning_layout(I, &l);
    char* b = (char*)base; BinView v;
    v.vals_sorted = (uint32_t*)(b + l.vals_sorted);
    v.ckpt = (float*)(b + l.ckpt); v.slot_tile = (uint32_t*)(b + l.slot_tile);
    v.ckpt_list = (uint32_t*)(b + l.ckpt_list);
    v.work_counter = (uint32_t*)(b + l.work_counter);
    return v;
}
ImageView tgs_image_view(void* base, int W, int H) {
    TgsImageLayout l; tgs_image_layout(W, H, &l);
    char* b = (char*)base; ImageView v;
    v.final_T = (float*)(b + l.final_T); v.n_contrib = (uint32_t*)(b + l.n_contrib);
    v.depth_raw = (float*)(b + l.depth_raw); v.color_acc = (float*)(b + l.color_acc);
    v.ranges = (uint2*)(b + l.ranges); v.count = (uint32_t*)(b + l.count);
    return v;
}

// ---------------------------------------------------------------------------- validation
static int check_inputs(const TgsSettings* s, const TgsGaussians* g) {
    if (!s || !g) { tgs_set_error("NULL settings / gaussians"); return TGS_EINVAL; }
    if (s->image_width <= 0 || s->image_height <= 0) { tgs_set_error("bad image size %dx%d", s->image_width, s->image_height); return TGS_EINVAL; }
    if (s->image_width > TGS_BIN_SCATTER_MAX_TX * TGS_TILE || s->image_height > 65535 * TGS_TILE) {
        tgs_set_error("image too large (width <= %d, height <= %d)", TGS_BIN_SCATTER_MAX_TX * TGS_TILE, 65535 * TGS_TILE); return TGS_EINVAL; }
    if (g->N < 0) { tgs_set_error("negative N"); return TGS_EINVAL; }
    if (!s->viewmatrix || !s->projmatrix || !s->bg) { tgs_set_error("viewmatrix / projmatrix / bg must be non-NULL"); return TGS_EINVAL; }
    if (g->N > 0) {
        if (!g->means3D || !g->opacities) { tgs_set_error("means3D / opacities must be non-NULL"); return TGS_EINVAL; }
        if ((g->shs == nullptr) == (g->colors_precomp == nullptr)) {
            tgs_set_error("Please provide exactly one of either SHs or precomputed colors!"); return TGS_EINVAL; }
        bool sr = g->scales != nullptr && g->rotations != nullptr;
        bool any_sr = g->scales != nullptr || g->rotations != nullptr;
        if ((sr == (g->cov3D_precomp != nullptr)) || (any_sr && !sr)) {
            tgs_set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!"); return TGS_EINVAL; }
        if (g->shs) {
            if (s->sh_degree < 0 || s->sh_degree > 3) { tgs_set_error("sh_degree %d out of range 0..3", s->sh_degree); return TGS_EINVAL; }
            int need = (s->sh_degree + 1) * (s->sh_degree + 1);
            if (s->sh_coeffs < need || s->sh_coeffs > 16) { tgs_set_error("sh_coeffs %d incompatible with degree %d", s->sh_coeffs, s->sh_degree); return TGS_EINVAL; }
            if (!s->campos) { tgs_set_error("campos must be non-NULL with SHs"); return TGS_EINVAL; }
        }
    }
    return 0;
}

static cudaEvent_t count_event() {
    static thread_local cudaEvent_t ev = nullptr;
    if (!ev) { if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) ev = nullptr; }
    return ev;
}

static uint32_t* pinned_word() {
    static thread_local uint32_t* p = nullptr;
    if (!p) { if (cudaHostAlloc((void**)&p, 64, cudaHostAllocDefault) != cudaSuccess) p = nullptr; }
    return p;
}

// ---------------------------------------------------------- deferred count (TgsSettings.defer_count)
// A forward that does not wait for num_rendered parks the D2H copy of the count in a pinned slot with an event and hands
// out a ticket; the backward (possibly on autograd's thread: the table is global) redeems it.
namespace {
struct CountTicket { uint32_t* host = nullptr; cudaEvent_t ev = nullptr; int state = 0; int64_t value = 0; };   // 0 free, 1 pending, 2 resolved
constexpr int kTickets = 8;
CountTicket g_tickets[kTickets];
int g_ticket_next = 0;
std::mutex g_ticket_mu;
}  // namespace
static int ticket_issue(const uint32_t* count_dev, cudaStream_t st, int64_t* ticket) {
    std::lock_guard<std::mutex> lk(g_ticket_mu);
    const int k = g_ticket_next;
    g_ticket_next = (g_ticket_next + 1) % kTickets;
    CountTicket& t = g_tickets[k];
    if (!t.host) TGS_CUDA(cudaHostAlloc((void**)&t.host, 64, cudaHostAllocDefault));
    if (!t.ev) TGS_CUDA(cudaEventCreateWithFlags(&t.ev, cudaEventDisableTiming));
    if (t.state == 1) TGS_CUDA(cudaEventSynchronize(t.ev));      // an unredeemed ticket 8 forwards old: its copy must land first
    TGS_CUDA(cudaMemcpyAsync(t.host, count_dev, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    TGS_CUDA(cudaEventRecord(t.ev, st));
    t.state = 1;
    *ticket = -(int64_t)(k + 1);
    return 0;
}
extern "C" int tgs_forward_resolve(int64_t ticket, int64_t capacity, int64_t* num_rendered_out) {
    if (ticket >= 0) { if (num_rendered_out) *num_rendered_out = ticket; return 0; }
    const int k = (int)(-ticket) - 1;
    if (k < 0 || k >= kTickets || !num_rendered_out) { tgs_set_error("tgs_forward_resolve: bad ticket"); return TGS_EINVAL; }
    std::lock_guard<std::mutex> lk(g_ticket_mu);
    CountTicket& t = g_tickets[k];
    if (t.state == 0) { tgs_set_error("tgs_forward_resolve: ticket was never issued or has expired"); return TGS_ESTATE; }
    if (t.state == 1) {
        TGS_CUDA(cudaEventSynchronize(t.ev));
        if (t.host[1]) { tgs_set_error("num_rendered does not fit 32 bits (more than 4,294,967,295 tile instances)"); return TGS_EINVAL; }
        t.value = (int64_t)t.host[0];
        t.state = 2;
    }
    *num_rendered_out = t.value;
    if (capacity >= 0 && t.value > capacity) {
        tgs_set_error("deferred speculative forward overflowed: %lld instances rendered, buffers sized for %lld (rendered_hint); "
                      "its outputs are invalid -- redo the step with a larger hint or without defer_count",
                      (long long)t.value, (long long)capacity);
        return TGS_EOVERFLOW;
    }
    return 0;
}

// ------------------------------------------------------------------------------ C ABI
extern "C" int tgs_abi_version(void) { return TGS_ABI_VERSION; }
extern "C" const char* tgs_last_error(void) { return g_err; }
extern "C" void tgs_launch_counts(uint64_t* own, uint64_t* cub) {
    if (own) *own = g_own.load();
    if (cub) *cub = g_cub.load();
}

extern "C" int tgs_mark_visible(int32_t N, const float* means3D, const float* viewmatrix, uint8_t* present, void* stream) {
    if (N < 0 || (N > 0 && (!means3D || !viewmatrix || !present))) { tgs_set_error("tgs_mark_visible: bad arguments"); return TGS_EINVAL; }
    return tgs_launch_mark_visible(N, means3D, viewmatrix, present, (cudaStream_t)stream);
}

extern "C" int tgs_forward(const TgsSettings* s, const TgsGaussians* g, tgs_alloc_fn alloc, void* user,
                           float* out_color, float* out_depth, float* out_alpha, int32_t* radii,
                           const float* touch_target, float* residual_out,
                           TgsSaved* saved, void* stream) {
    int rc = check_inputs(s, g);
    if (rc) return rc;
    if (!alloc || !saved || !out_color || !out_depth || !out_alpha || (g->N > 0 && !radii)) {
        tgs_set_error("tgs_forward: NULL output / allocator"); return TGS_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const TgsCam cam = tgs_make_cam(s);
    const int N = g->N, T = cam.Tx * cam.Ty;

    TgsGeomLayout gl; tgs_geom_layout(N, &gl);
    TgsImageLayout il; tgs_image_layout(cam.W, cam.H, &il);
    void* geom = alloc(user, TGS_BUF_GEOM, gl.total);
    void* image = alloc(user, TGS_BUF_IMAGE, il.total);
    if (!geom || !image) { tgs_set_error("allocator returned NULL"); return TGS_ENOMEM; }
    GeomView gv = tgs_geom_view(geom, N);
    ImageView iv = tgs_image_view(image, cam.W, cam.H);

    if ((touch_target == nullptr) != (residual_out == nullptr)) { tgs_set_error("touch_target and residual_out go together"); return TGS_EINVAL; }
    int64_t I = 0, cap = 0;
    void* binning = nullptr;
    void* temp = nullptr;
    uint32_t* hp = pinned_word();
    if (!hp) { tgs_set_error("cudaHostAlloc failed"); return TGS_ENOMEM; }
    auto tail = [&](int64_t count, int64_t capacity, bool spec) -> int {      // scatter + render for `capacity` slots
        TgsBinningLayout bl; tgs_binning_layout(capacity, &bl);
        binning = alloc(user, TGS_BUF_BINNING, bl.total);
        if (!binning) { tgs_set_error("allocator returned NULL"); return TGS_ENOMEM; }
        BinView bv = tgs_bin_view(binning, capacity);
        int r = tgs_bin_scatter(gv, bv, N, count, capacity, spec, cam.Tx, cam.Ty, cam.row0, cam.row1, temp, iv.ranges, iv.count, st); if (r) return r;
        if (s->debug) TGS_CUDA(cudaStreamSynchronize(st));
        return tgs_launch_render_fwd(cam, s, gv.records, bv, iv, capacity, out_color, out_depth, out_alpha, touch_target, residual_out, st);
    };
    auto read_count = [&]() -> int {                    // hp[0] = num_rendered, hp[1] = 32-bit overflow flag
        if (hp[1]) { tgs_set_error("num_rendered does not fit 32 bits (more than 4,294,967,295 tile instances)"); return TGS_EINVAL; }
        I = (int64_t)hp[0];
        return 0;
    };
    if (N > 0) {
        const size_t tb = tgs_bin_temp_bytes(N, cam.Tx, cam.Ty);
        temp = alloc(user, TGS_BUF_TEMP, tb);
        if (!temp) { tgs_set_error("allocator returned NULL"); return TGS_ENOMEM; }
        rc = tgs_launch_preprocess(cam, s, g, gv, radii, st); if (rc) return rc;
        rc = tgs_depth_order(gv, N, st); if (rc) return rc;
        rc = tgs_bin_count(gv, N, cam.Tx, cam.Ty, cam.row0, cam.row1, temp, iv.ranges, iv.count, st); if (rc) return rc;
        if (s->rendered_hint > 0 && s->defer_count) {
            // DEFERRED: the whole forward is enqueued for `hint` slots and the host never waits; the count travels to a
            // pinned slot behind an event and is checked when the backward redeems the ticket
            cap = s->rendered_hint;
            rc = ticket_issue(iv.count, st, &I); if (rc) return rc;
            rc = tail(cap, cap, true); if (rc) return rc;
            saved->geom = geom; saved->binning = binning; saved->image = image; saved->num_rendered = I; saved->capacity = cap;
            return 0;
        }
        TGS_CUDA(cudaMemcpyAsync(hp, iv.count, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (s->rendered_hint > 0) {
            // SPECULATIVE: enqueue scatter + render for `hint` slots, THEN wait for the count (event recorded
            // right after the count kernels: the GPU keeps working on the speculative tail while the host wakes up)
            cudaEvent_t ev = count_event();
            if (!ev) { tgs_set_error("cudaEventCreate failed"); return TGS_ENOMEM; }
            TGS_CUDA(cudaEventRecord(ev, st));
            cap = s->rendered_hint;
            rc = tail(cap, cap, true); if (rc) return rc;
            TGS_CUDA(cudaEventSynchronize(ev));
            rc = read_count(); if (rc) return rc;
            if (I > cap) { cap = I; rc = tail(I, I, false); if (rc) return rc; }   // hint too small: exact re-run
        } else {
            TGS_CUDA(cudaStreamSynchronize(st));   // the one host sync of the forward
            rc = read_count(); if (rc) return rc;
            cap = I;
            rc = tail(I, I, false); if (rc) return rc;
        }
    } else {
        rc = tgs_bin_count(gv, 0, cam.Tx, cam.Ty, cam.row0, cam.row1, nullptr, iv.ranges, iv.count, st); if (rc) return rc;
        rc = tail(0, 0, false); if (rc) return rc;
    }
    saved->geom = geom; saved->binning = binning; saved->image = image; saved->num_rendered = I; saved->capacity = cap;
    return 0;
}

extern "C" size_t tgs_screen_grad_bytes(int32_t N, int32_t contrib_flags) {
    if (N <= 0) return 0;
    if (!contrib_flags) return sizeof(float) * TGS_NGRAD * (size_t)N;
    return TGS_SCREEN_GRAD_FLAG_OFFSET(N) + ((size_t)N + 127u) / 128u * 128u;
}

extern "C" int tgs_backward_render(const TgsSettings* s, const TgsGaussians* g, const TgsSaved* saved,
                                   const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                                   const TgsTouch* touch, float* residual_out, float* screen_grads, void* stream) {
    int rc = check_inputs(s, g);
    if (rc) return rc;
    if (!saved || !saved->geom || !saved->binning || !saved->image) { tgs_set_error("tgs_backward_render: saved buffers missing"); return TGS_ESTATE; }
    if (!dL_dcolor || (g->N > 0 && !screen_grads)) { tgs_set_error("tgs_backward_render: NULL gradient buffers"); return TGS_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const TgsCam cam = tgs_make_cam(s);
    int64_t I = saved->num_rendered;
    if (I < 0) { rc = tgs_forward_resolve(I, saved->capacity, &I); if (rc) return rc; }     // deferred forward: redeem the count
    BinView bv = tgs_bin_view(saved->binning, saved->capacity > 0 ? saved->capacity : I);
    ImageView iv = tgs_image_view(saved->image, cam.W, cam.H);
    if (g->N > 0) TGS_CUDA(cudaMemsetAsync(screen_grads, 0, tgs_screen_grad_bytes(g->N, s->contrib_flags), st));
    GeomView gvb = tgs_geom_view(saved->geom, g->N);
    uint8_t* flags = (s->contrib_flags && g->N > 0) ? (uint8_t*)screen_grads + TGS_SCREEN_GRAD_FLAG_OFFSET(g->N) : nullptr;
    return tgs_launch_render_bwd(cam, s, gvb.records, bv, iv, I, dL_dcolor, dL_ddepth, dL_dalpha, touch, residual_out,
                                 screen_grads, flags, st);
}

extern "C" int tgs_backward_preprocess(const TgsSettings* s, const TgsGaussians* g, const TgsSaved* saved,
                                       const int32_t* radii, const float* screen_grads, const TgsGrads* grads, void* stream) {
    int rc = check_inputs(s, g);
    if (rc) return rc;
    if (g->N == 0) return 0;
    if (!saved || !saved->geom) { tgs_set_error("tgs_backward_preprocess: saved buffers missing"); return TGS_ESTATE; }
    if (!radii || !screen_grads || !grads || !grads->dmeans2D || !grads->dmeans3D || !grads->dopacity) {
        tgs_set_error("tgs_backward_preprocess: NULL gradient buffers"); return TGS_EINVAL; }
    if (g->shs && !grads->dshs) { tgs_set_error("dshs required when shs given"); return TGS_EINVAL; }
    if (g->colors_precomp && !grads->dcolors) { tgs_set_error("dcolors required when colors_precomp given"); return TGS_EINVAL; }
    if (g->scales && (!grads->dscales || !grads->drotations)) { tgs_set_error("dscales/drotations required"); return TGS_EINVAL; }
    if (g->cov3D_precomp && !grads->dcov3D) { tgs_set_error("dcov3D required when cov3D_precomp given"); return TGS_EINVAL; }
    const TgsCam cam = tgs_make_cam(s);
    GeomView gv = tgs_geom_view(saved->geom, g->N);
    return tgs_launch_preprocess_bwd(cam, s, g, gv, radii, screen_grads, nullptr, nullptr, 0, s->contrib_flags != 0, grads,
                                     (cudaStream_t)stream);
}

extern "C" int tgs_backward_preprocess_gather(const TgsSettings* s, const TgsGaussians* g, const TgsSaved* saved,
                                              const int32_t* radii, const float* const* peer_screen_grads_host,
                                              const int32_t* peer_tile_rows_host, int32_t world,
                                              const TgsGrads* grads, void* stream) {
    int rc = check_inputs(s, g);
    if (rc) return rc;
    if (g->N == 0) return 0;
    if (!saved || !saved->geom) { tgs_set_error("tgs_backward_preprocess_gather: saved buffers missing"); return TGS_ESTATE; }
    if (world < 1 || world > TGS_MAX_PEERS || !peer_screen_grads_host || !peer_tile_rows_host) {
        tgs_set_error("tgs_backward_preprocess_gather: world must be 1..%d with peer pointers and bands", TGS_MAX_PEERS); return TGS_EINVAL; }
    for (int r = 0; r < world; ++r)
        if (!peer_screen_grads_host[r]) { tgs_set_error("tgs_backward_preprocess_gather: peer %d pointer is NULL", r); return TGS_EINVAL; }
    if (!radii || !grads || !grads->dmeans2D || !grads->dmeans3D || !grads->dopacity) {
        tgs_set_error("tgs_backward_preprocess_gather: NULL gradient buffers"); return TGS_EINVAL; }
    if (g->shs && !grads->dshs) { tgs_set_error("dshs required when shs given"); return TGS_EINVAL; }
    if (g->colors_precomp && !grads->dcolors) { tgs_set_error("dcolors required when colors_precomp given"); return TGS_EINVAL; }
    if (g->scales && (!grads->dscales || !grads->drotations)) { tgs_set_error("dscales/drotations required"); return TGS_EINVAL; }
    if (g->cov3D_precomp && !grads->dcov3D) { tgs_set_error("dcov3D required when cov3D_precomp given"); return TGS_EINVAL; }
    const TgsCam cam = tgs_make_cam(s);
    GeomView gv = tgs_geom_view(saved->geom, g->N);
    return tgs_launch_preprocess_bwd(cam, s, g, gv, radii, nullptr, peer_screen_grads_host, peer_tile_rows_host, world,
                                     s->contrib_flags != 0, grads, (cudaStream_t)stream);
}

extern "C" int tgs_backward(const TgsSettings* s, const TgsGaussians* g, const TgsSaved* saved, const int32_t* radii,
                            const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                            const TgsTouch* touch, float* residual_out, float* screen_grads,
                            const TgsGrads* grads, void* stream) {
    int rc = tgs_backward_render(s, g, saved, dL_dcolor, dL_ddepth, dL_dalpha, touch, residual_out, screen_grads, stream);
    if (rc) return rc;
    return tgs_backward_preprocess(s, g, saved, radii, screen_grads, grads, stream);
}

extern "C" int tgs_touch_loss_scale(const float* target, int64_t num_pixels, float mult, float norm, float* scale_out, void* stream) {
    if (!scale_out || (norm <= 0.0f && num_pixels > 0 && !target)) { tgs_set_error("tgs_touch_loss_scale: bad arguments"); return TGS_EINVAL; }
    return tgs_launch_loss_scale(target, num_pixels, mult, norm, scale_out, (cudaStream_t)stream);
}

extern "C" int tgs_touch_loss_value(const float* residual, const float* weight, int32_t W, int32_t H, int32_t row_begin,
                                    int32_t row_end, int32_t mode, const float* scale, double* acc, float* loss_out,
                                    void* stream) {
    if (!residual || !scale || !acc || !loss_out || W <= 0 || H <= 0) { tgs_set_error("tgs_touch_loss_value: bad arguments"); return TGS_EINVAL; }
    if (mode != TGS_LOSS_NONE && mode != TGS_LOSS_L1 && mode != TGS_LOSS_L2) { tgs_set_error("tgs_touch_loss_value: bad mode %d", mode); return TGS_EINVAL; }
    int r0 = 0, r1 = H;
    if (row_end > row_begin) { r0 = row_begin < 0 ? 0 : row_begin; r1 = row_end > H ? H : row_end; }
    return tgs_launch_touch_loss_value(residual, weight, (int64_t)r0 * W, (int64_t)r1 * W, mode, scale, acc, loss_out,
                                       (cudaStream_t)stream);
}
