// k10_filters.cu -- K10: GPU pre-filters hqdn3d (denoise) and yadif (deinterlace) behind include/b2enc_filters.h.
//
// SURVEY.md 8f row N4.  In the reference these run inside the libavfilter graph in front of sws_scale / x264
// (av_encode.c:451-560, recommended chain "hqdn3d,yadif", :35).  Bit-exact against oracle/b2o_filters.c (which restates the
// published algorithms; libavfilter itself is not in the image -> parity unpinned beyond that).
//
// hqdn3d is three nested first-order recursions with a table-driven, similarity-weighted coefficient: along the row, down the
// column, and over time.  The recursions are non-linear (table lookup on the difference), so no parallel scan applies; but they
// separate: K10a walks every ROW sequentially (one thread per row, all rows and planes in parallel), K10b walks every COLUMN
// (vertical + temporal recursion, coalesced across the warp).  yadif (K10c) is stateless per pixel given prev / cur / next.
// Bound: K10b and K10c HBM (one read + one write of the picture and of the 16-bit state); K10a latency of the row walk.
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <deque>
#include <string>
#include <vector>
#include "b2_common.cuh"
#include "../../include/b2enc_filters.h"

namespace {

constexpr int LUT_BITS = 4, LUT_HALF = 256 << LUT_BITS;

__device__ __forceinline__ int lowpass(int prev, int cur, const int32_t *__restrict__ ct) { return cur + __ldg(&ct[LUT_HALF + ((prev - cur) >> (8 - LUT_BITS))]); }
__device__ __forceinline__ int load8(int px) { return (px << 8) + 127; }

// K10a: horizontal recursion, one thread per row: hbuf[y][x] (8.8 fixed point)
__global__ void k10a_hqdn3d_rows_kernel(const uint8_t *__restrict__ src, int w, int h, uint16_t *__restrict__ hbuf, const int32_t *__restrict__ spatial)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    const uint8_t *row = src + (size_t)y * w;
    uint16_t *out = hbuf + (size_t)y * w;
    int pixel_ant = load8(row[0]);
    for (int x = 0; x < w; x++) {
        if (x) pixel_ant = lowpass(pixel_ant, load8(row[x]), spatial);
        out[x] = (uint16_t)pixel_ant;
    }
}
// K10b: vertical + temporal recursion, one thread per column
__global__ void k10b_hqdn3d_cols_kernel(const uint8_t *__restrict__ src, const uint16_t *__restrict__ hbuf, uint8_t *__restrict__ dst, int w, int h,
                                        uint16_t *__restrict__ frame_ant, int first, const int32_t *__restrict__ spatial,
                                        const int32_t *__restrict__ temporal)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    int v = 0;
    for (int y = 0; y < h; y++) {
        const size_t i = (size_t)y * w + x;
        const int hv = hbuf[i];
        v = y ? lowpass(v, hv, spatial) : hv;
        const int prev = first ? load8(src[i]) : frame_ant[i];
        int t = lowpass(prev, v, temporal);
        t = min(max(t, 0), 65535);
        frame_ant[i] = (uint16_t)t;
        dst[i] = (uint8_t)min((t + 128) >> 8, 255);
    }
}

__device__ __forceinline__ int px_at(const uint8_t *__restrict__ p, int w, int h, int x, int y)
{
    x = min(max(x, 0), w - 1); y = min(max(y, 0), h - 1);
    return p[(size_t)y * w + x];
}
// K10c: yadif mode 0, one thread per pixel
__global__ void k10c_yadif_kernel(const uint8_t *__restrict__ prev, const uint8_t *__restrict__ cur, const uint8_t *__restrict__ next,
                                  uint8_t *__restrict__ dst, int w, int h, int tff)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int parity = tff ^ 1;
    if (!((y ^ parity) & 1)) { dst[(size_t)y * w + x] = cur[(size_t)y * w + x]; return; }
    const uint8_t *prev2 = prev, *next2 = cur;
    const int ym = y ? y - 1 : y + 1, yp = y + 1 < h ? y + 1 : y - 1;
    const bool limited = y == 1 || y + 2 == h || y == 0 || y + 1 == h;
#define CUR(xx, yy) px_at(cur, w, h, xx, yy)
    const int c = CUR(x, ym), e = CUR(x, yp);
    const int p2 = px_at(prev2, w, h, x, y), n2 = px_at(next2, w, h, x, y);
    const int d = (p2 + n2) >> 1;
    const int td0 = abs(p2 - n2);
    const int td1 = (abs(px_at(prev, w, h, x, ym) - c) + abs(px_at(prev, w, h, x, yp) - e)) >> 1;
    const int td2 = (abs(px_at(next, w, h, x, ym) - c) + abs(px_at(next, w, h, x, yp) - e)) >> 1;
    int diff = max(max(td0 >> 1, td1), td2);
    int spatial_pred = (c + e) >> 1;
    int spatial_score = abs(CUR(x - 1, ym) - CUR(x - 1, yp)) + abs(c - e) + abs(CUR(x + 1, ym) - CUR(x + 1, yp)) - 1;
#pragma unroll
    for (int dir = -1; dir <= 1; dir += 2)
#pragma unroll
        for (int k = 1; k <= 2; k++) {
            const int j = dir * k;
            const int score = abs(CUR(x - 1 + j, ym) - CUR(x - 1 - j, yp)) + abs(CUR(x + j, ym) - CUR(x - j, yp)) +
                              abs(CUR(x + 1 + j, ym) - CUR(x + 1 - j, yp));
            if (score >= spatial_score) break;
            spatial_score = score;
            spatial_pred = (CUR(x + j, ym) + CUR(x - j, yp)) >> 1;
        }
    if (!limited) {
        const int b = (px_at(prev2, w, h, x, y - 2) + px_at(next2, w, h, x, y - 2)) >> 1;
        const int f = (px_at(prev2, w, h, x, y + 2) + px_at(next2, w, h, x, y + 2)) >> 1;
        const int mx = max(max(d - e, d - c), min(b - c, f - e));
        const int mn = min(min(d - e, d - c), max(b - c, f - e));
        diff = max(max(diff, mn), -mx);
    }
    spatial_pred = min(max(spatial_pred, d - diff), d + diff);
    dst[(size_t)y * w + x] = (uint8_t)spatial_pred;
#undef CUR
}

// ---- graph ----------------------------------------------------------------------------------------------------------------
struct DFrame { uint8_t *p[3] = {}; int64_t pts = 0; int tff = 1; };

struct Stage {
    int type = 0;                              // 1 hqdn3d, 2 yadif
    // hqdn3d
    int32_t *ct[4] = {};                       // luma spatial, luma temporal, chroma spatial, chroma temporal
    uint16_t *frame_ant[3] = {}, *hbuf = nullptr;
    bool first = true;
    // yadif
    int parity = -1;
    DFrame *prev = nullptr, *cur = nullptr;
};

}  // namespace

struct b2_filter_graph {
    int device, w, h, fmt, np;
    int pw[3], ph[3];
    cudaStream_t st = nullptr;
    uint8_t *h_stage = nullptr;                // pinned staging for one frame
    size_t frame_bytes = 0;
    std::vector<Stage> stages;
    std::deque<DFrame *> ready;
    std::vector<DFrame *> pool;
};

namespace {

DFrame *frame_get(b2_filter_graph *g)
{
    if (!g->pool.empty()) { DFrame *f = g->pool.back(); g->pool.pop_back(); return f; }
    DFrame *f = new DFrame();
    for (int p = 0; p < g->np; p++)
        if (cudaMalloc(&f->p[p], (size_t)g->pw[p] * g->ph[p]) != cudaSuccess) { fprintf(stderr, "b2enc: filter frame allocation failed\n"); return nullptr; }
    return f;
}
void frame_put(b2_filter_graph *g, DFrame *f) { if (f) g->pool.push_back(f); }

void hqdn3d_coefs(double dist25, std::vector<int32_t> &ct)
{
    ct.assign(2 * LUT_HALF, 0);
    if (dist25 <= 0) return;
    const double d = dist25 < 252.0 ? dist25 : 252.0;
    const double gamma = log(0.25) / log(1.0 - d / 255.0 - 0.00001);
    for (int i = -LUT_HALF; i < LUT_HALF; i++) {
        const double f = (double)((i * (1 << (9 - LUT_BITS))) + (1 << (8 - LUT_BITS)) - 1) / 512.0;
        double simil = 1.0 - fabs(f) / 255.0;
        if (simil < 0) simil = 0;
        ct[LUT_HALF + i] = (int32_t)lrint(pow(simil, gamma) * 256.0 * f);
    }
}

int yadif_run(b2_filter_graph *g, DFrame *prev, DFrame *cur, DFrame *next, DFrame *out, int parity)
{
    const int tff = parity < 0 ? cur->tff : (parity ^ 1);
    for (int p = 0; p < g->np; p++) {
        dim3 grid((g->pw[p] + 127) / 128, g->ph[p]);
        k10c_yadif_kernel<<<grid, 128, 0, g->st>>>(prev->p[p], cur->p[p], next->p[p], out->p[p], g->pw[p], g->ph[p], tff);
    }
    B2_CUDA_OK(cudaGetLastError());
    out->pts = cur->pts; out->tff = cur->tff;
    return 0;
}

// run frame f (or a flush when f == nullptr) through stages [s0, end); finished frames go to the ready queue
int run_from(b2_filter_graph *g, size_t s0, DFrame *f)
{
    for (size_t si = s0; si < g->stages.size(); si++) {
        Stage &s = g->stages[si];
        if (s.type == 1) {
            if (!f) continue;                                  // nothing held back
            DFrame *o = frame_get(g);
            if (!o) return -1;
            for (int p = 0; p < g->np; p++) {
                const int w = g->pw[p], h = g->ph[p];
                const int32_t *sp = s.ct[p ? 2 : 0], *tp = s.ct[p ? 3 : 1];
                k10a_hqdn3d_rows_kernel<<<(h + 63) / 64, 64, 0, g->st>>>(f->p[p], w, h, s.hbuf, sp);
                k10b_hqdn3d_cols_kernel<<<(w + 127) / 128, 128, 0, g->st>>>(f->p[p], s.hbuf, o->p[p], w, h, s.frame_ant[p], s.first ? 1 : 0, sp, tp);
            }
            B2_CUDA_OK(cudaGetLastError());
            s.first = false;
            o->pts = f->pts; o->tff = f->tff;
            frame_put(g, f);
            f = o;
        } else {                                               // yadif: one frame of delay (it needs the next picture)
            if (f) {
                DFrame *out = nullptr;
                if (s.cur) {
                    out = frame_get(g);
                    if (!out || yadif_run(g, s.prev ? s.prev : s.cur, s.cur, f, out, s.parity)) return -1;
                }
                frame_put(g, s.prev);
                s.prev = s.cur; s.cur = f;
                f = out;
                if (!f) return 0;                              // first frame: nothing to hand on yet
            } else if (s.cur) {                                // flush: the last picture is its own successor
                DFrame *out = frame_get(g);
                if (!out || yadif_run(g, s.prev ? s.prev : s.cur, s.cur, s.cur, out, s.parity)) return -1;
                frame_put(g, s.prev); frame_put(g, s.cur);
                s.prev = s.cur = nullptr;
                if (run_from(g, si + 1, out)) return -1;       // the released frame first, then the flush of the later stages
            }
        }
    }
    if (f) g->ready.push_back(f);
    return 0;
}

}  // namespace

extern "C" b2_filter_graph_t *b2_filter_graph_create(int width, int height, int fmt, const char *filters, int device)
{
    if (width < 8 || height < 8 || (fmt != B2_FMT_YUV420P && fmt != B2_FMT_YUV422P && fmt != B2_FMT_YUV411P)) {
        fprintf(stderr, "b2enc: b2_filter_graph_create: planar yuv420p / yuv422p / yuv411p, at least 8x8\n");
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        fprintf(stderr, "b2enc: no CUDA device; the filter stage has no CPU fallback\n");
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { fprintf(stderr, "b2enc: cannot select device %d\n", device); return nullptr; }
    b2_filter_graph *g = new b2_filter_graph();
    g->device = device; g->w = width; g->h = height; g->fmt = fmt;
    int rb[3], rows[3];
    g->np = b2_fmt_layout(fmt, width, height, rb, rows);
    for (int p = 0; p < 3; p++) { g->pw[p] = rb[p]; g->ph[p] = rows[p]; g->frame_bytes += (size_t)rb[p] * rows[p]; }
    // parse "name[=a[:b[:c[:d]]]]" joined by ','
    std::string spec = filters ? filters : "";
    size_t pos = 0;
    while (pos < spec.size()) {
        size_t end = spec.find(',', pos);
        if (end == std::string::npos) end = spec.size();
        std::string item = spec.substr(pos, end - pos);
        pos = end + 1;
        if (item.empty()) continue;
        std::string name = item.substr(0, item.find('=')), args = item.find('=') == std::string::npos ? "" : item.substr(item.find('=') + 1);
        double a[4] = {-1, -1, -1, -1};
        int na = 0;
        for (size_t q = 0; q < args.size() && na < 4;) {
            size_t e2 = args.find(':', q);
            if (e2 == std::string::npos) e2 = args.size();
            a[na++] = atof(args.substr(q, e2 - q).c_str());
            q = e2 + 1;
        }
        Stage s;
        if (name == "hqdn3d") {
            s.type = 1;
            const double ls = na > 0 && a[0] >= 0 ? a[0] : 4.0, cs = na > 1 && a[1] >= 0 ? a[1] : 3.0 * ls / 4.0;
            const double lt = na > 2 && a[2] >= 0 ? a[2] : 6.0 * ls / 4.0, ctmp = na > 3 && a[3] >= 0 ? a[3] : (ls > 0 ? lt * cs / ls : 0.0);
            const double str[4] = {ls, lt, cs, ctmp};
            bool ok = cudaMalloc(&s.hbuf, (size_t)width * height * 2) == cudaSuccess;
            for (int k = 0; k < 4 && ok; k++) {
                std::vector<int32_t> ct;
                hqdn3d_coefs(str[k], ct);
                ok = cudaMalloc(&s.ct[k], ct.size() * 4) == cudaSuccess && cudaMemcpy(s.ct[k], ct.data(), ct.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
            }
            for (int p = 0; p < g->np && ok; p++) ok = cudaMalloc(&s.frame_ant[p], (size_t)g->pw[p] * g->ph[p] * 2) == cudaSuccess;
            if (!ok) { fprintf(stderr, "b2enc: hqdn3d allocation failed\n"); g->stages.push_back(s); b2_filter_graph_free(g); return nullptr; }
        } else if (name == "yadif") {
            s.type = 2;
            if (na > 0 && a[0] != 0) { fprintf(stderr, "b2enc: yadif: only mode 0 is implemented\n"); b2_filter_graph_free(g); return nullptr; }
            s.parity = na > 1 ? (int)a[1] : -1;
        } else {
            fprintf(stderr, "b2enc: unknown filter '%s' (hqdn3d and yadif are implemented)\n", name.c_str());   // av_encode.c:499-502
            b2_filter_graph_free(g);
            return nullptr;
        }
        g->stages.push_back(s);
    }
    if (cudaStreamCreate(&g->st) != cudaSuccess || cudaHostAlloc(&g->h_stage, g->frame_bytes, cudaHostAllocDefault) != cudaSuccess) {
        fprintf(stderr, "b2enc: filter graph allocation failed\n");
        b2_filter_graph_free(g);
        return nullptr;
    }
    return g;
}

extern "C" void b2_filter_graph_free(b2_filter_graph_t *g)
{
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->st) cudaStreamSynchronize(g->st);
    auto kill = [](DFrame *f) { if (f) { for (int p = 0; p < 3; p++) cudaFree(f->p[p]); delete f; } };
    for (auto &s : g->stages) {
        for (int k = 0; k < 4; k++) cudaFree(s.ct[k]);
        for (int p = 0; p < 3; p++) cudaFree(s.frame_ant[p]);
        cudaFree(s.hbuf);
        kill(s.prev); kill(s.cur);
    }
    for (auto f : g->ready) kill(f);
    for (auto f : g->pool) kill(f);
    cudaFreeHost(g->h_stage);
    if (g->st) cudaStreamDestroy(g->st);
    delete g;
}

extern "C" int b2_filter_add_frame(b2_filter_graph_t *g, const uint8_t *const src[3], const int stride[3], int64_t pts, int top_field_first)
{
    if (!g || !src) return -1;
    cudaSetDevice(g->device);
    DFrame *f = frame_get(g);
    if (!f) return -1;
    uint8_t *hp = g->h_stage;
    for (int p = 0; p < g->np; p++) {
        for (int y = 0; y < g->ph[p]; y++) memcpy(hp + (size_t)y * g->pw[p], src[p] + (size_t)y * stride[p], (size_t)g->pw[p]);
        B2_CUDA_OK(cudaMemcpyAsync(f->p[p], hp, (size_t)g->pw[p] * g->ph[p], cudaMemcpyHostToDevice, g->st));
        hp += (size_t)g->pw[p] * g->ph[p];
    }
    f->pts = pts; f->tff = top_field_first ? 1 : 0;
    if (run_from(g, 0, f)) return -1;
    B2_CUDA_OK(cudaStreamSynchronize(g->st));                  // the caller may free the source now (av_encode.c reuses its AVFrame)
    return 0;
}

extern "C" int b2_filter_poll_frame(b2_filter_graph_t *g) { return g ? (int)g->ready.size() : 0; }

extern "C" int b2_filter_get_frame(b2_filter_graph_t *g, uint8_t *const dst[3], const int stride[3], int64_t *pts)
{
    if (!g || !dst) return -1;
    if (g->ready.empty()) return 0;
    cudaSetDevice(g->device);
    DFrame *f = g->ready.front();
    g->ready.pop_front();
    for (int p = 0; p < g->np; p++)
        B2_CUDA_OK(cudaMemcpy2DAsync(dst[p], (size_t)stride[p], f->p[p], (size_t)g->pw[p], (size_t)g->pw[p], (size_t)g->ph[p], cudaMemcpyDeviceToHost, g->st));
    B2_CUDA_OK(cudaStreamSynchronize(g->st));
    if (pts) *pts = f->pts;
    frame_put(g, f);
    return 1;
}

extern "C" int b2_filter_flush(b2_filter_graph_t *g)
{
    if (!g) return -1;
    cudaSetDevice(g->device);
    if (run_from(g, 0, nullptr)) return -1;
    B2_CUDA_OK(cudaStreamSynchronize(g->st));
    return 0;
}
