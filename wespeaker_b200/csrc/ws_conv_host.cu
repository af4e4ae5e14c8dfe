// Host side of the fused conv operator: tap generation (incl. parity planes for strided convs), TMA tensor-map
// construction, tile-shape selection and the launch closures for the tcgen05 and FFMA kernels.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>

#include "../../include/wespeaker_b200.h"
#include "ws_host.h"

namespace ws {

static thread_local std::string g_err;
void set_err(const std::string& msg) { g_err = msg; }
static thread_local std::string g_op_label;
static thread_local double g_op_flops = 0.0;
static thread_local bool g_op_has = false;
void set_op_label(const std::string& name, double flops) { g_op_label = name; g_op_flops = flops; g_op_has = true; }
bool take_op_label(std::string* name, double* flops) {
    if (!g_op_has) return false;
    *name = g_op_label; *flops = g_op_flops; g_op_has = false;
    return true;
}
const std::string& get_err() { return g_err; }

static thread_local bool g_plan_check = false;
bool plan_check_mode() { return g_plan_check; }
PlanCheckScope::PlanCheckScope(bool on) : prev(g_plan_check) { g_plan_check = on; }
PlanCheckScope::~PlanCheckScope() { g_plan_check = prev; }
static thread_local std::string g_op_trace;
static thread_local bool g_op_trace_has = false;
void set_op_trace(const std::string& json) {
    if (!g_plan_check) return;
    g_op_trace = json; g_op_trace_has = true;
}
bool take_op_trace(std::string* json) {
    if (!g_op_trace_has) return false;
    *json = g_op_trace; g_op_trace_has = false;
    return true;
}
static std::string jp(const void* p) { char b[32]; snprintf(b, sizeof b, "%llu", (unsigned long long)p); return b; }
static std::string jv(const char* name, long long v) { char b[96]; snprintf(b, sizeof b, "\"%s\":%lld", name, v); return b; }
static std::string view_json(const View& v) {
    return "{\"p\":" + jp(v.p) + "," + jv("B", v.B) + "," + jv("F", v.F) + "," + jv("T", v.T) + "," + jv("C", v.C) + "," + jv("ld", v.ld) + "}";
}
static std::string conv_trace(const ConvSpec& s) {
    std::string j = "{\"kind\":\"conv\"," + jv("es", ws_esize(s.dt)) + ",\"src\":[";
    for (int i = 0; i < s.nsrc; ++i) {
        const WsSrc& v = s.src[i];
        j += std::string(i ? "," : "") + "{\"p\":" + jp(v.ptr) + "," + jv("B", v.B) + "," + jv("F", v.F) + "," + jv("T", v.T) + "," + jv("C", v.C) + "," +
             jv("sB", v.sB) + "," + jv("sF", v.sF) + "," + jv("sT", v.sT) + "}";
    }
    j += "],\"taps\":[";
    for (size_t i = 0; i < s.taps.size(); ++i) {
        const WsTap& t = s.taps[i];
        j += std::string(i ? "," : "") + "[" + std::to_string(t.src) + "," + std::to_string(t.c0) + "," + std::to_string(t.dt) + "," + std::to_string(t.df) + "," +
             std::to_string(t.wk) + "," + std::to_string(t.nch) + "]";
    }
    const WsEpi& e = s.epi;
    j += "],\"W\":" + jp(s.W) + "," + jv("Ktot", s.Ktot) + "," + jv("Cout", s.Cout) + "," + jv("B", s.B) + "," + jv("F", s.F) + "," + jv("T", s.T) +
         ",\"bias\":" + jp(e.bias) + ",\"rowbias\":" + jp(e.rowbias) + "," + jv("rowbias_ld", e.rowbias_ld) + "," + jv("act1", e.act1) + ",\"scale\":" + jp(e.scale) +
         ",\"shift\":" + jp(e.shift) + ",\"gate\":" + jp(e.gate) + "," + jv("gate_ld", e.gate_ld) + "," + jv("gate_seg", e.gate_seg) + "," + jv("gate_nseg", e.gate_nseg) + ",\"res\":" + jp(e.res) + "," + jv("res_ld", e.res_ld) + "," + jv("act2", e.act2) + ",\"out\":" + jp(e.out) +
         "," + jv("out_ld", e.out_ld) + ",\"out2\":" + jp(e.out2) + "," + jv("out2_ld", e.out2_ld) + ",\"add2\":" + jp(e.add2) + "," + jv("add2_ld", e.add2_ld) +
         ",\"colsum\":" + jp(e.colsum) + "}";
    return j;
}

void* plan_check_alloc(size_t bytes) {
    static thread_local unsigned long long next = 0x7000000000ull;   // far from anything a host allocator returns
    void* p = (void*)next;
    next += (bytes + 255) & ~255ull;
    if (bytes == 0) next += 256;
    return p;
}

void fill_epi_out(WsEpi& e, const View& out) {
    e.out = out.p;
    e.out_lo = out.plo;
    e.out_ld = out.ld;
    e.dtype = out.dt;
    e.FT = out.F * out.T;
    e.T = out.T;
}

static int find_or_add_src(ConvSpec& s, const WsSrc& v) {
    for (int i = 0; i < s.nsrc; ++i)
        if (s.src[i].ptr == v.ptr && s.src[i].sT == v.sT && s.src[i].sF == v.sF && s.src[i].T == v.T &&
            s.src[i].F == v.F && s.src[i].C == v.C)
            return i;
    if (s.nsrc >= WS_MAX_SRC) return -1;
    s.src[s.nsrc] = v;
    return s.nsrc++;
}

static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

int add_conv_taps(ConvSpec& spec, const View& x, int kf, int kt, int dil_f, int dil_t, int pad_f, int pad_t,
                  int stride_f, int stride_t, int wk0, int* Fo, int* To) {
    const int es = ws_esize(x.dt);
    *Fo = (x.F + 2 * pad_f - dil_f * (kf - 1) - 1) / stride_f + 1;
    *To = (x.T + 2 * pad_t - dil_t * (kt - 1) - 1) / stride_t + 1;
    const long long sT = x.ld, sF = (long long)x.T * x.ld, sB = (long long)x.F * x.T * x.ld;
    for (int jf = 0; jf < kf; ++jf)
        for (int jt = 0; jt < kt; ++jt) {
            const int offf = jf * dil_f - pad_f, offt = jt * dil_t - pad_t;
            const int pf = ((offf % stride_f) + stride_f) % stride_f, pt = ((offt % stride_t) + stride_t) % stride_t;
            WsSrc v;
            v.ptr = (const char*)x.p + ((size_t)pf * sF + (size_t)pt * sT) * es;
            v.ptr_lo = x.plo ? (const char*)x.plo + ((size_t)pf * sF + (size_t)pt * sT) * es : nullptr;
            v.B = x.B;
            v.F = (x.F - pf + stride_f - 1) / stride_f;
            v.T = (x.T - pt + stride_t - 1) / stride_t;
            v.C = x.C;
            v.sB = sB;
            v.sF = sF * stride_f;
            v.sT = sT * stride_t;
            const int si = find_or_add_src(spec, v);
            if (si < 0) return -1;
            WsTap tap;
            tap.src = si;
            tap.c0 = 0;
            tap.df = floordiv(offf - pf, stride_f);
            tap.dt = floordiv(offt - pt, stride_t);
            tap.wk = wk0 + (jf * kt + jt) * x.C;
            tap.nch = x.C;
            spec.taps.push_back(tap);
        }
    return kf * kt * x.C;
}

// ------------------------------------------------------------------------------------------------ TMA maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

static bool encode_map(CUtensorMap* m, int dt, const void* ptr, int rank, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes) {
    if (plan_check_mode()) {
        // no driver: check what cuTensorMapEncodeTiled would check (CUDA driver API, tensor map object management)
        const int es = ws_esize(dt);
        char buf[200];
        const char* why = nullptr;
        if (((unsigned long long)ptr & 15ull) != 0) why = "global address not 16-byte aligned";
        if (rank < 1 || rank > 5) why = "rank outside 1..5";
        for (int i = 0; i < rank && !why; ++i) {
            if (dims[i] < 1 || dims[i] > (1ull << 32)) why = "dimension outside 1..2^32";
            else if (box[i] < 1 || box[i] > 256) why = "box extent outside 1..256";
            else if (i + 1 < rank && (strides_bytes[i] % 16 != 0 || strides_bytes[i] >= (1ull << 40))) why = "stride not a multiple of 16 bytes (or >= 2^40)";
        }
        if (!why) {
            const unsigned inner = box[0] * (unsigned)es;
            if (swizzle_bytes != 0 && inner > (unsigned)swizzle_bytes) why = "inner box extent exceeds the swizzle span";
            else if (inner % 16 != 0) why = "inner box extent not a multiple of 16 bytes";
        }
        if (why) {
            snprintf(buf, sizeof buf, "tensor map check failed: %s (rank %d dims [%llu %llu ..] box [%u %u ..] sw %d)", why, rank,
                     (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0, swizzle_bytes);
            set_err(buf);
            return false;
        }
        memset(m, 0, sizeof(*m));
        return true;
    }
    EncodeTiledFn fn = get_encode();
    if (fn == nullptr) {
        set_err("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
        return false;
    }
    const CUtensorMapDataType t = dt == WS_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                               : (dt == WS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                                : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                     : (swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_32B));
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, t, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[256];
        snprintf(buf, sizeof buf,
                 "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u] sw %d", (int)r,
                 rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                 (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                 rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, swizzle_bytes);
        set_err(buf);
        return false;
    }
    return true;
}

static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

static bool build_tc(const ConvSpec& s, WsTcParams* p) {
    memset(p, 0, sizeof(*p));
    const int es = ws_esize(s.dt);
    p->kind = s.dt == WS_F32 ? 0 : 1;
    if (s.Cout % 16 != 0) { set_err("tc conv: Cout must be a multiple of 16"); return false; }
    if ((int)s.taps.size() > WS_MAX_TAPS) { set_err("tc conv: too many taps"); return false; }
    int minb = 1 << 30;
    for (const WsTap& t : s.taps) minb = std::min(minb, t.nch * es);
    p->bk_bytes = minb >= 128 ? 128 : (minb >= 64 ? 64 : 32);
    if (minb < 32) { set_err("tc conv: fewer than 32 bytes of channels per tap"); return false; }
    const int bk_elems = p->bk_bytes / es;
    p->ntaps = (int)s.taps.size();
    p->nk_total = 0;
    for (int i = 0; i < p->ntaps; ++i) {
        const WsTap& t = s.taps[i];
        WsTcTap& o = p->taps[i];
        o.map = t.src; o.c0 = t.c0; o.dt = t.dt; o.df = t.df; o.wk = t.wk;
        o.nkb = (t.nch + bk_elems - 1) / bk_elems;
        if (t.nch % bk_elems != 0 && t.c0 + t.nch != s.src[t.src].C) {
            set_err("tc conv: ragged k-block needs the tap to end at the source view's last channel");
            return false;
        }
        p->nk_total += o.nkb;
    }
    // ---- output tile shape
    int B = s.B, F = s.F, T = s.T;
    bool flat = s.dense_pointwise;
    int bt = 128, bf = 1, bb = 1;
    const char* env_bt = getenv("WS_TILE_BT");  // tuning knob: minimum time-extent of a non-flat tile
    const int min_bt = env_bt ? atoi(env_bt) : 1;
    if (flat) {
        T = s.B * s.F * s.T; F = 1; B = 1;
    } else {
        long long best = -1;
        for (int t = 1; t <= 128; t <<= 1)
            for (int f = 1; t * f <= 128; f <<= 1) {
                const int b = 128 / (t * f);
                if (F == 1 && f != 1) continue;
                if (t < min_bt && t < 128 && F == 1) continue;
                const long long cost = (long long)((T + t - 1) / t) * t * ((F + f - 1) / f) * f * ((B + b - 1) / b) * b;
                if (best < 0 || cost < best || (cost == best && t > bt)) { best = cost; bt = t; bf = f; bb = b; }
            }
    }
    p->bt_log2 = ilog2(bt); p->bf_log2 = ilog2(bf); p->bb_log2 = ilog2(bb);
    p->tiles_t = (T + bt - 1) / bt; p->tiles_f = (F + bf - 1) / bf; p->tiles_b = (B + bb - 1) / bb;
    p->B = B; p->F = F; p->T = T;
    p->bn = s.Cout % 128 == 0 ? 128 : (s.Cout % 64 == 0 ? 64 : (s.Cout % 32 == 0 ? 32 : 16));
    p->tiles_n = s.Cout / p->bn;
    if ((p->bn * p->bk_bytes) % 1024 != 0) { set_err("tc conv: weight tile not a multiple of the 1024-B swizzle atom"); return false; }
    const int stage_bytes = (128 + p->bn) * p->bk_bytes;
    int ns = (96 * 1024) / stage_bytes;
    p->nstages = ns < 2 ? 2 : (ns > WS_TC_MAX_STAGES ? WS_TC_MAX_STAGES : ns);
    const uint32_t fmt = s.dt == WS_F32 ? 2u : (s.dt == WS_BF16 ? 1u : 0u);
    // cute::UMMA::InstrDescriptor: c_format f32 @4, a_format @7, b_format @10, K-major both, N>>3 @17, M>>4 @24
    p->idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p->bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // ---- tensor maps
    for (int i = 0; i < s.nsrc; ++i) {
        const WsSrc& v = s.src[i];
        cuuint64_t dims[4], str[3];
        cuuint32_t box[4];
        if (flat) {
            dims[0] = (cuuint64_t)v.C; dims[1] = (cuuint64_t)T; dims[2] = 1; dims[3] = 1;
            str[0] = (cuuint64_t)v.sT * es; str[1] = (cuuint64_t)T * v.sT * es; str[2] = str[1];
        } else {
            dims[0] = (cuuint64_t)v.C; dims[1] = (cuuint64_t)v.T; dims[2] = (cuuint64_t)v.F; dims[3] = (cuuint64_t)v.B;
            str[0] = (cuuint64_t)v.sT * es; str[1] = (cuuint64_t)v.sF * es; str[2] = (cuuint64_t)v.sB * es;
        }
        box[0] = (cuuint32_t)bk_elems; box[1] = (cuuint32_t)bt; box[2] = (cuuint32_t)bf; box[3] = (cuuint32_t)bb;
        if (!encode_map(&p->amap[i], s.dt, v.ptr, 4, dims, str, box, p->bk_bytes)) return false;
    }
    for (int i = s.nsrc; i < WS_MAX_SRC; ++i) p->amap[i] = p->amap[0];
    {
        cuuint64_t dims[2] = {(cuuint64_t)s.Ktot, (cuuint64_t)s.Cout};
        cuuint64_t str[1] = {(cuuint64_t)s.Ktot * es};
        cuuint32_t box[2] = {(cuuint32_t)bk_elems, (cuuint32_t)p->bn};
        if (!encode_map(&p->wmap, s.dt, s.W, 2, dims, str, box, p->bk_bytes)) return false;
    }
    p->epi = s.epi;
    return true;
}


// ---- v2 (persistent, TMA-store epilogue).  Returns false *without* error if the spec needs the v1 kernel.
static bool build_tc2(const ConvSpec& s, WsTc2Params* q, bool* unsupported, int cl = 1) {
    const bool pair = cl >= 2;   // cl: CTAs per cluster (1 = ws_gemm_tc2; 2 / 4 = ws_gemm_tc3, see there)
    *unsupported = false;
    const WsEpi& e = s.epi;
    if ((e.res != nullptr && e.out2 != nullptr) || s.Cout % 32 != 0) { *unsupported = true; return false; }
    WsTcParams v1;
    if (!build_tc(s, &v1)) return false;
    memset(q, 0, sizeof(*q));
    const int es = ws_esize(s.dt);
    for (int i = 0; i < WS_MAX_SRC; ++i) q->amap[i] = v1.amap[i];
    memcpy(q->taps, v1.taps, sizeof(v1.taps));
    q->ntaps = v1.ntaps; q->nk_total = v1.nk_total; q->bk_bytes = v1.bk_bytes;
    q->bt_log2 = v1.bt_log2; q->bf_log2 = v1.bf_log2; q->bb_log2 = v1.bb_log2;
    q->tiles_t = v1.tiles_t; q->tiles_f = v1.tiles_f; q->tiles_b = v1.tiles_b;
    q->B = v1.B; q->F = v1.F; q->T = v1.T; q->kind = v1.kind;
    q->has_out2 = e.out2 != nullptr;
    q->has_epin = (e.res != nullptr) || (e.out2 != nullptr);
    q->nsplit = s.split ? 3 : 1;
    if (s.split) {
        if (s.dt != WS_F32 || s.W_lo == nullptr || e.out_lo == nullptr || (q->has_out2 && e.out2_lo == nullptr)) {
            set_err("tc conv: 3xTF32 needs fp32 activations with lo twins for W / out / out2"); return false;
        }
        for (int i = 0; i < s.nsrc; ++i)
            if (s.src[i].ptr_lo == nullptr) { set_err("tc conv: 3xTF32 source without a lo twin"); return false; }
    }
    q->nout = (s.split ? 2 : 1) * (q->has_out2 ? 2 : 1);
    const int nstage_bufs = q->nout + (q->has_epin ? 1 : 0);
    const int budget = ws_tc2_max_smem() - 1024;
    // widest N tile whose staging + >=3 ring stages fit (>=2 accepted as a last resort)
    int bn = 0, nst = 0;
    const int cands[4] = {256, 128, 64, 32};
    const char* env_bn = getenv("WS_TC2_MAX_BN");
    const int max_bn = env_bn ? atoi(env_bn) : 256;
    const char* env_st = getenv("WS_TC2_MAX_STAGES");
    const int max_st = env_st ? atoi(env_st) : WS_TC_MAX_STAGES;
    for (int pass = 0; pass < 2 && bn == 0; ++pass)
        for (int ci = 0; ci < 4; ++ci) {
            const int c = cands[ci];
            if (s.Cout % c != 0 || c > max_bn) continue;
            if (c == 256 && es == 4) continue;  // fp32 staging of a 128x256 tile would not leave room for the ring
            if (pair && c < 128) continue;      // cta_group::2: each CTA stages bn/2 weight rows
            const int wrows = pair ? c / 2 : c;
            if ((wrows * q->bk_bytes) % 1024 != 0) continue;
            const int staging = nstage_bufs * 128 * c * es + 3 * c * 4;
            const int stage_bytes = (128 + wrows) * q->bk_bytes;
            const int n = (budget - staging) / stage_bytes;
            if (n >= (pass == 0 ? 3 : 2)) { bn = c; nst = n > max_st ? max_st : n; break; }
        }
    if (bn == 0) { *unsupported = true; return false; }
    q->bn = bn; q->nstages = nst;
    q->tiles_n = s.Cout / bn;
    const int mtiles = q->tiles_t * q->tiles_f * q->tiles_b;
    q->num_tiles = q->tiles_n * ((mtiles + cl - 1) / cl);
    const uint32_t fmt = s.dt == WS_F32 ? 2u : (s.dt == WS_BF16 ? 1u : 0u);
    q->idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((pair ? 256 : 128) >> 4) << 24);
    q->panel_bytes = bn * es >= 128 ? 128 : bn * es;
    const int panel_cols = q->panel_bytes / es;
    const int bk_elems = q->bk_bytes / es;
    {   // weights map with the v2 N tile
        cuuint64_t dims[2] = {(cuuint64_t)s.Ktot, (cuuint64_t)s.Cout};
        cuuint64_t str[1] = {(cuuint64_t)s.Ktot * es};
        cuuint32_t box[2] = {(cuuint32_t)bk_elems, (cuuint32_t)(bn / cl)};   // the rows ONE CTA loads per k-block
        if (!encode_map(&q->wmap, s.dt, s.W, 2, dims, str, box, q->bk_bytes)) return false;
        q->wmap_lo = q->wmap;
        if (s.split && !encode_map(&q->wmap_lo, s.dt, s.W_lo, 2, dims, str, box, q->bk_bytes)) return false;
    }
    for (int i = 0; i < WS_MAX_SRC; ++i) q->amap_lo[i] = q->amap[i];
    if (s.split) {   // lo twins of the activation maps: same geometry, different base pointer
        ConvSpec lo = s;
        for (int i = 0; i < lo.nsrc; ++i) lo.src[i].ptr = lo.src[i].ptr_lo;
        WsTcParams v1lo;
        if (!build_tc(lo, &v1lo)) return false;
        for (int i = 0; i < WS_MAX_SRC; ++i) q->amap_lo[i] = v1lo.amap[i];
    }
    // output / epilogue-input maps: same tile geometry as the A maps over the output positions
    const bool flat = s.dense_pointwise;
    auto out_map = [&](CUtensorMap* m, const void* ptr, long long ld) -> bool {
        cuuint64_t dims[4], str[3];
        cuuint32_t box[4] = {(cuuint32_t)panel_cols, (cuuint32_t)(1 << q->bt_log2), (cuuint32_t)(1 << q->bf_log2),
                             (cuuint32_t)(1 << q->bb_log2)};
        if (flat) {
            dims[0] = (cuuint64_t)s.Cout; dims[1] = (cuuint64_t)q->T; dims[2] = 1; dims[3] = 1;
            str[0] = (cuuint64_t)ld * es; str[1] = (cuuint64_t)q->T * ld * es; str[2] = str[1];
        } else {
            dims[0] = (cuuint64_t)s.Cout; dims[1] = (cuuint64_t)s.T; dims[2] = (cuuint64_t)s.F; dims[3] = (cuuint64_t)s.B;
            str[0] = (cuuint64_t)ld * es; str[1] = (cuuint64_t)s.T * ld * es; str[2] = (cuuint64_t)s.F * s.T * ld * es;
        }
        return encode_map(m, s.dt, ptr, 4, dims, str, box, q->panel_bytes);
    };
    int oi = 0;
    if (!out_map(&q->omap[oi++], e.out, e.out_ld)) return false;
    if (s.split && !out_map(&q->omap[oi++], e.out_lo, e.out_ld)) return false;
    if (q->has_out2) {
        if (!out_map(&q->omap[oi++], e.out2, e.out2_ld)) return false;
        if (s.split && !out_map(&q->omap[oi++], e.out2_lo, e.out2_ld)) return false;
        if (!out_map(&q->imap, e.add2, e.add2_ld)) return false;
    } else if (q->has_epin) {
        if (!out_map(&q->imap, e.res, e.res_ld)) return false;
    }
    for (; oi < 4; ++oi) q->omap[oi] = q->omap[0];
    if (!q->has_epin) q->imap = q->omap[0];
    const int g_num_sms = ws_num_sms();
    q->cl = cl;
    if (pair) {
        int ncl = ws_tc3_max_clusters(cl);
        if (ncl <= 0) ncl = g_num_sms / cl;
        q->grid = cl * (q->num_tiles < ncl ? q->num_tiles : ncl);
    } else {
        q->grid = q->num_tiles < g_num_sms ? q->num_tiles : g_num_sms;
    }
    q->smem_bytes = nst * (128 + (pair ? bn / 2 : bn)) * q->bk_bytes + nstage_bufs * 128 * bn * es + 3 * bn * 4 + 1024;
    q->epi = e;
    const char* env_sh = getenv("WS_TC2_SHIFT_TEST");
    q->dbg_shift = env_sh ? atoi(env_sh) : -1;
    { const char* eg = getenv("WS_EPI_GENERIC"); q->epi_generic = eg != nullptr && atoi(eg) != 0; }
    return true;
}

static bool build_simt(const ConvSpec& s, WsSimtParams* p) {
    memset(p, 0, sizeof(*p));
    if ((int)s.taps.size() > WS_MAX_TAPS) { set_err("conv: too many taps"); return false; }
    if (s.Cout % 4 != 0) { set_err("conv: Cout must be a multiple of 4"); return false; }
    for (int i = 0; i < s.nsrc; ++i) p->src[i] = s.src[i];
    p->ntaps = (int)s.taps.size();
    for (int i = 0; i < p->ntaps; ++i) p->taps[i] = s.taps[i];
    p->W = s.W; p->Ktot = s.Ktot; p->Cout = s.Cout;
    p->B = s.B; p->F = s.F; p->T = s.T;
    p->dtype = s.dt;
    p->epi = s.epi;
    return true;
}

static void label_conv(const char* kern, const ConvSpec& spec) {
    const double pos = (double)spec.B * spec.F * spec.T;
    char buf[160];
    snprintf(buf, sizeof buf, "%s pos=%lld K=%d N=%d", kern, (long long)pos, spec.Ktot, spec.Cout);
    set_op_label(buf, 2.0 * pos * spec.Ktot * spec.Cout);
}
bool make_conv_op(const ConvSpec& spec, int use_tc, Op* out) {
    if (plan_check_mode()) set_op_trace(conv_trace(spec));
    const char* env_mp = getenv("WS_TC3_MIN_POS");
    const long long min_pos = env_mp ? atoll(env_mp) : 148LL * 128;
    if (use_tc >= 3 && spec.Cout % 128 == 0 && (long long)spec.B * spec.F * spec.T >= min_pos) {
        // big layers: CTA pairs (cta_group::2) halve the weight-operand traffic per CTA
        auto q = std::make_shared<WsTc2Params>();
        bool unsupported = false;
        // ... and clusters of two pairs share the weight tile by TMA multicast (WS_TC3_CL=2 keeps single pairs)
        static const int cl_env = getenv("WS_TC3_CL") ? atoi(getenv("WS_TC3_CL")) : 2;
        const int cl = (cl_env == 4 && (long long)spec.B * spec.F * spec.T >= 2LL * min_pos) ? 4 : 2;
        if (build_tc2(spec, q.get(), &unsupported, cl)) {
            *out = [q](cudaStream_t s) { return ws_tc3_launch(q.get(), s); };
            {
                char kn[64];
                snprintf(kn, sizeof kn, "conv_tc3 cl=%d grid=%d bn=%d st=%d", q->cl, q->grid, q->bn, q->nstages);
                label_conv(kn, spec);
            }
            return true;
        }
        if (!unsupported) return false;
    }
    if (use_tc >= 2) {
        auto q = std::make_shared<WsTc2Params>();
        bool unsupported = false;
        if (build_tc2(spec, q.get(), &unsupported)) {
            *out = [q](cudaStream_t s) { return ws_tc2_launch(q.get(), s); };
            label_conv("conv_tc2", spec);
            return true;
        }
        if (!unsupported) return false;
    }
    if (spec.epi.colsum != nullptr) { set_err("conv colsum: only the persistent tensor-core kernels (use_tc >= 2) produce column sums"); return false; }
    if (use_tc) {
        auto p = std::make_shared<WsTcParams>();
        if (!build_tc(spec, p.get())) return false;
        *out = [p](cudaStream_t s) { return ws_tc_launch(p.get(), s); };
        label_conv("conv_tc1", spec);
    } else {
        auto p = std::make_shared<WsSimtParams>();
        if (!build_simt(spec, p.get())) return false;
        *out = [p](cudaStream_t s) { return ws_simt_launch(p.get(), s); };
        label_conv("conv_simt", spec);
    }
    return true;
}

bool make_res2_op(const View& x, const View& out, const void* W7, const float* bias, const float* scale,
                  const float* shift, int w8, int dil, Op* op, bool* unsupported, const int* lens) {
    *unsupported = false;
    // T > 256: time tiles with a 32-row halo per side cover the chain's receptive field (7 * dil rows) up to dilation 4
    if (x.dt == WS_F32 || (w8 != 64 && w8 != 128) || (x.T > 256 && dil > 4) || x.F != 1 || dil < 1 || dil > 7 || getenv("WS_NO_RES2_FUSED")) {
        *unsupported = true;
        return false;
    }
    auto q = std::make_shared<WsRes2Params>();
    memset(q.get(), 0, sizeof(WsRes2Params));
    auto act_map = [&](CUtensorMap* m, const View& v) {
        cuuint64_t dims[3] = {(cuuint64_t)v.C, (cuuint64_t)v.T, (cuuint64_t)v.B};
        cuuint64_t str[2] = {(cuuint64_t)v.ld * 2, (cuuint64_t)v.T * v.ld * 2};
        cuuint32_t box[3] = {64, 128, 1};
        return encode_map(m, v.dt, v.p, 3, dims, str, box, 128);
    };
    if (!act_map(&q->xmap, x) || !act_map(&q->omap, out)) return false;
    {
        cuuint64_t dims[3] = {(cuuint64_t)out.C, (cuuint64_t)out.T, (cuuint64_t)out.B};
        cuuint64_t str[2] = {(cuuint64_t)out.ld * 2, (cuuint64_t)out.T * out.ld * 2};
        cuuint32_t box[3] = {64, 64, 1};
        if (!encode_map(&q->omap64, out.dt, out.p, 3, dims, str, box, 128)) return false;
    }
    q->ntile = x.T <= 256 ? 1 : (x.T + 191) / 192;
    {
        cuuint64_t dims[2] = {(cuuint64_t)(3 * w8), (cuuint64_t)(7 * w8)};
        cuuint64_t str[1] = {(cuuint64_t)(3 * w8) * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)w8};
        if (!encode_map(&q->wmap, x.dt, W7, 2, dims, str, box, 128)) return false;
    }
    q->x = x.p; q->ld = x.ld; q->bias = bias; q->scale = scale; q->shift = shift; q->lens = lens;
    q->B = x.B; q->T = x.T; q->w8 = w8; q->dil = dil; q->dtype = x.dt;
    const uint32_t fmt = x.dt == WS_BF16 ? 1u : 0u;
    q->idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(w8 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int g_num_sms = ws_num_sms();
    const int npan = w8 / 64;
    q->smem_bytes = npan * 272 * 128 + 4 * w8 * 128 + npan * 256 * 128 + 1024;
    // 64-wide groups need ~100 KB: two 256-thread CTAs per SM, so that one utterance's MMAs run under another's epilogue
    q->ew = (w8 == 64 && q->smem_bytes <= 112 * 1024 && !getenv("WS_RES2_EW8")) ? 4 : 8;
    const int slots = g_num_sms * (q->ew == 4 ? 2 : 1);
    const long long units = (long long)x.B * q->ntile;
    q->grid = (int)(units < slots ? units : slots);
    *op = [q](cudaStream_t s) { return ws_res2_launch(q.get(), s); };
    if (plan_check_mode())
        set_op_trace("{\"kind\":\"res2_fused\",\"es\":2,\"x\":" + view_json(x) + ",\"out\":" + view_json(out) + ",\"W7\":" + jp(W7) + ",\"bias\":" + jp(bias) +
                     ",\"scale\":" + jp(scale) + ",\"shift\":" + jp(shift) + "," + jv("w8", w8) + "," + jv("dil", dil) + ",\"lens\":" + jp(lens) + "}");
    {
        char buf[160];
        snprintf(buf, sizeof buf, "res2_fused B=%d T=%d C=%d", x.B, x.T, x.C);
        set_op_label(buf, 2.0 * x.B * x.T * 7.0 * 3.0 * (x.C / 8.0) * (x.C / 8.0));
    }
    return true;
}

bool make_astp_op(const View& x, const View& h, const void* W2, float* stats, Op* op, bool* unsupported, const int* lens) {
    *unsupported = false;
    if (x.dt == WS_F32 || h.dt != x.dt || x.C % 128 != 0 || h.C != 128 || h.ld != 128 || x.F != 1 || h.F != 1 || h.T != x.T ||
        h.B != x.B || getenv("WS_NO_ASTP_FUSED")) {
        *unsupported = true;
        return false;
    }
    auto q = std::make_shared<WsAstpParams>();
    memset(q.get(), 0, sizeof(WsAstpParams));
    {
        cuuint64_t dims[3] = {128, (cuuint64_t)h.T, (cuuint64_t)h.B};
        cuuint64_t str[2] = {128 * 2, (cuuint64_t)h.T * 128 * 2};
        cuuint32_t box[3] = {64, 256, 1};
        if (!encode_map(&q->hmap, h.dt, h.p, 3, dims, str, box, 128)) return false;
    }
    {
        cuuint64_t dims[2] = {128, (cuuint64_t)x.C};
        cuuint64_t str[1] = {128 * 2};
        cuuint32_t box[2] = {64, 128};
        if (!encode_map(&q->wmap, x.dt, W2, 2, dims, str, box, 128)) return false;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)x.C, (cuuint64_t)x.T, (cuuint64_t)x.B};
        cuuint64_t str[2] = {(cuuint64_t)x.ld * 2, (cuuint64_t)x.T * x.ld * 2};
        cuuint32_t box[3] = {128, 128, 1};
        if (!encode_map(&q->xmap, x.dt, x.p, 3, dims, str, box, 0)) return false;
    }
    q->x = x.p; q->x_ld = x.ld; q->out = stats; q->B = x.B; q->T = x.T; q->C = x.C; q->dtype = x.dt; q->lens = lens;
    // channel blocks per unit: the largest g whose static round-robin over the SMs stays within 8 % of the best balance
    const int nblk = x.C / 128, sms = ws_num_sms();
    double best = 0.0;
    int cand[6] = {12, 6, 4, 3, 2, 1};
    auto eff = [&](int g) {
        const long long units = (long long)x.B * (nblk / g);
        const long long grid = units < sms ? units : sms;
        return (double)units / (double)(((units + grid - 1) / grid) * grid);
    };
    for (int g : cand) if (nblk % g == 0 && eff(g) > best) best = eff(g);
    q->g = 1;
    for (int g : cand) if (nblk % g == 0 && eff(g) >= 0.92 * best) { q->g = g; break; }
    if (const char* eg = getenv("WS_ASTP_G")) { const int g = atoi(eg); if (g > 0 && nblk % g == 0) q->g = g; }
    const long long units = (long long)x.B * (nblk / q->g);
    q->grid = (int)(units < sms ? units : sms);
    if (getenv("WS_ASTP_PROF")) {   // tuning aid: synchronous launch + per-role wait-cycle summary on stderr
        long long* prof = nullptr;
        if (cudaMalloc((void**)&prof, (size_t)q->grid * 16 * 8) != cudaSuccess) { set_err("astp: prof buffer"); return false; }
        q->prof = prof;
        *op = [q](cudaStream_t s) -> const char* {
            cudaMemsetAsync(q->prof, 0, (size_t)q->grid * 16 * 8, s);
            const char* m = ws_astp_launch(q.get(), s);
            if (m) return m;
            cudaStreamSynchronize(s);
            std::vector<long long> h((size_t)q->grid * 16);
            cudaMemcpy(h.data(), q->prof, h.size() * 8, cudaMemcpyDeviceToHost);
            static const char* names[11] = {"prod.wait_hempty", "prod.wait_wempty", "mma.wait_hfull", "mma.wait_wfull", "mma.wait_tempty",
                                            "mma.tiles", "epi0.wait_tfull", "epi0.body", "epi1.wait_tfull", "epi1.body", "total"};
            fprintf(stderr, "[astp prof] grid %d g %d:", q->grid, q->g);
            for (int k = 0; k < 11; ++k) {
                double acc = 0;
                for (int c = 0; c < q->grid; ++c) acc += (double)h[(size_t)c * 16 + k];
                fprintf(stderr, " %s=%.0f", names[k], acc / q->grid);
            }
            fprintf(stderr, "\n");
            return nullptr;
        };
        set_op_label("astp_fused (profiled)", 0.0);
        return true;
    }
    *op = [q](cudaStream_t s) { return ws_astp_launch(q.get(), s); };
    if (plan_check_mode())
        set_op_trace("{\"kind\":\"astp_fused\",\"es\":2,\"x\":" + view_json(x) + ",\"h\":" + view_json(h) + ",\"W2\":" + jp(W2) + ",\"stats\":" + jp(stats) +
                     ",\"lens\":" + jp(lens) + "}");
    {
        char buf[160];
        snprintf(buf, sizeof buf, "astp_fused B=%d T=%d C=%d g=%d grid=%d", x.B, x.T, x.C, q->g, q->grid);
        set_op_label(buf, 2.0 * x.B * x.T * 128.0 * x.C);
    }
    return true;
}

bool make_conv3x3_op(const View& x, const View& out, const void* W, const float* bias, const View* res, int relu,
                     Op* op, bool* unsupported, int stride_f, int stride_t, const int* lens) {
    *unsupported = false;
    const int Cin = x.C, Cout = out.C, Tin = x.T, Fin = x.F, B = x.B;
    const int F = (Fin + 2 - 3) / stride_f + 1, T = (Tin + 2 - 3) / stride_t + 1;   // output extents
    auto chan_ok = [](int c) { return c == 32 || c == 64 || c == 128; };
    if (x.dt == WS_F32 || out.dt != x.dt || !chan_ok(Cin) || !chan_ok(Cout) || out.B != B || out.F != F || out.T != T ||
        (stride_f != 1 && stride_f != 2) || (stride_t != 1 && stride_t != 2) || T < 1 || F < 1 || (x.ld * 2) % 16 != 0 || (out.ld * 2) % 16 != 0 || (res && (res->ld * 2) % 16 != 0) ||
        getenv("WS_NO_CONV3X3")) {
        *unsupported = true;
        return false;
    }
    auto q = std::make_shared<WsC3Params>();
    memset(q.get(), 0, sizeof(WsC3Params));
    q->B = B; q->F = F; q->T = T; q->Cin = Cin; q->Cout = Cout; q->dtype = x.dt;
    q->sf = stride_f; q->st = stride_t;
    q->row_bytes = Cin * 2 >= 128 ? 128 : Cin * 2;
    q->kc = q->row_bytes / 2;
    q->npan = Cin / q->kc;
    q->N = Cout; q->n_nt = 1;
    // ---- geometry of one ring slot / one step: first candidate whose ring (>= 4 slots) fits in shared memory
    const int budget = ws_c3_max_smem();
    auto try_geom = [&](bool case_b, int tb, int n_mt) -> bool {
        q->case_b = case_b ? 1 : 0;
        q->tb = tb; q->n_mt = n_mt;
        q->n_tt = (T + tb - 1) / tb;
        q->P = tb + 2;
        if (case_b) {             // several short utterances side by side in one M tile (pitch P = T + 2)
            q->nb = std::min(B, 130 / q->P);   // nb * P - 2 <= 128
            q->single_box = 1;
            q->rows_loaded = q->nb * q->P;
        } else {                  // one utterance per slot, 1-2 M tiles along t
            q->nb = 1;
            q->single_box = q->P <= 256;
            q->rows_loaded = q->single_box ? q->P : 128 * n_mt + 2;
        }
        q->n_bg = (B + q->nb - 1) / q->nb;
        q->slot_rows = (std::max(q->rows_loaded, 128 * n_mt + 2) + 7) & ~7;
        q->sub_rows = 0;
        if (stride_t == 2) {      // even plane (tb rows) + shifted odd plane (tb + 1 rows), each in its own sub-slot
            if (case_b || n_mt != 1) return false;
            q->single_box = 1;
            q->sub_rows = (128 + 1 + 7) & ~7;
            q->slot_rows = 2 * q->sub_rows;
            q->rows_loaded = tb + (tb + 1);
        }
        if (2 * n_mt * q->N > 512) return false;
        // output panels (= TMA store units): 64 columns, or 32 when that is what it takes to give each of the two epilogue
        // warp sets its own unit
        q->panel_cols = std::min(64, q->N);
        if (q->panel_cols == 64 && n_mt * (q->N / 64) < 2) q->panel_cols = 32;
        q->panel_bytes = q->panel_cols * 2;
        q->stg_rows = case_b ? 136 : 128;
        const int slot_bytes = q->npan * q->slot_rows * q->row_bytes;
        const int wblk = q->N * q->row_bytes, nwblk = 9 * q->npan;
        const int stg_bytes = n_mt * (q->N / q->panel_cols) * q->stg_rows * q->panel_bytes;
        // three output staging buffers with a residual (its TMA load then leads its use by a full step), two without (stores
        // are never waited for), as long as a ring of >= 5 slots still fits beside them; else one
        for (int bufs = (res != nullptr && !getenv("WS_C3_NO_3BUF")) ? 3 : 2; bufs >= 1; --bufs) {
            const int fixed = bufs * stg_bytes + q->N * 4 + 2048 /* alignment slack */;
            q->w_resident = (budget - fixed - nwblk * wblk) / slot_bytes >= 5;
            q->w_stages = q->w_resident ? 0 : 3;
            int R = (budget - fixed - (q->w_resident ? nwblk : q->w_stages) * wblk) / slot_bytes;
            if (R > WS_C3_MAX_RING) R = WS_C3_MAX_RING;
            if (const char* er = getenv("WS_C3_RING")) R = std::min(R, atoi(er));
            if (R < (bufs >= 2 ? 5 : 4) + (stride_f - 1)) continue;
            q->R = R;
            q->stg_bufs = bufs;
            q->smem_bytes = R * slot_bytes + (q->w_resident ? nwblk : q->w_stages) * wblk + fixed;
            return true;
        }
        return false;
    };
    bool ok = false;
    const bool strided = stride_f != 1 || stride_t != 1;
    if (T + 2 <= 66 && !strided) ok = try_geom(true, T, 1);
    if (!ok && stride_t == 2) ok = try_geom(false, std::min(T, 128), 1);
    if (!ok && stride_t == 1 && T <= 256) ok = try_geom(false, T, (T + 127) / 128);
    if (!ok && stride_t == 1) {   // t tiles of 256 (2 M tiles) or 128 (1 M tile): fewer 128-row tiles first, 256 on a tie
        const bool pref256 = 2 * ((T + 255) / 256) <= (T + 127) / 128;
        ok = pref256 ? (try_geom(false, 256, 2) || try_geom(false, 128, 1)) : (try_geom(false, 128, 1) || try_geom(false, 256, 2));
    }
    if (!ok) { *unsupported = true; return false; }
    const uint32_t fmt = x.dt == WS_BF16 ? 1u : 0u;
    q->idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(q->N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    q->total_steps = q->n_nt * q->n_tt * q->n_bg * F;
    const int sms = ws_num_sms();
    q->grid = q->total_steps < sms ? q->total_steps : sms;
    // streamed weights (they do not fit beside the input ring) and a single channel tile: clusters of two CTAs can share the
    // weight stream by TMA multicast (each CTA fetches half of every block).  Measured on B200 (profiles/r02_conv3x3_layer3.md):
    // no gain - the step is bound by shared-memory bandwidth (8 KB of operand reads + 4 KB of weight-ring fill per MMA), which
    // multicast does not reduce - so it is opt-in (WS_C3_MC=1) and kept as a reproducible experiment.
    q->cl = (!q->w_resident && q->n_nt == 1 && q->grid >= 2 && q->N % 16 == 0 && getenv("WS_C3_MC")) ? 2 : 1;
    if (q->cl == 2) q->grid &= ~1;
    // ---- tensor maps
    if (stride_t == 2) {   // even / odd time planes of the input as two tensors with a doubled t stride
        cuuint64_t str[3] = {(cuuint64_t)x.ld * 4, (cuuint64_t)Tin * x.ld * 2, (cuuint64_t)Fin * Tin * x.ld * 2};
        cuuint64_t de[4] = {(cuuint64_t)Cin, (cuuint64_t)((Tin + 1) / 2), (cuuint64_t)Fin, (cuuint64_t)B};
        cuuint64_t dod[4] = {(cuuint64_t)Cin, (cuuint64_t)std::max(1, Tin / 2), (cuuint64_t)Fin, (cuuint64_t)B};
        cuuint32_t be[4] = {(cuuint32_t)q->kc, (cuuint32_t)q->tb, 1, 1}, bo[4] = {(cuuint32_t)q->kc, (cuuint32_t)(q->tb + 1), 1, 1};
        if (Tin < 2) { *unsupported = true; return false; }
        if (!encode_map(&q->amap, x.dt, x.p, 4, de, str, be, q->row_bytes)) return false;
        if (!encode_map(&q->amap_tail, x.dt, (const char*)x.p + (size_t)x.ld * 2, 4, dod, str, bo, q->row_bytes)) return false;
    } else {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)Tin, (cuuint64_t)Fin, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)x.ld * 2, (cuuint64_t)Tin * x.ld * 2, (cuuint64_t)Fin * Tin * x.ld * 2};
        cuuint32_t box[4] = {(cuuint32_t)q->kc, (cuuint32_t)(q->single_box ? q->P : 128), 1, (cuuint32_t)q->nb};
        if (!encode_map(&q->amap, x.dt, x.p, 4, dims, str, box, q->row_bytes)) return false;
        cuuint32_t boxt[4] = {(cuuint32_t)q->kc, 2, 1, 1};
        if (!encode_map(&q->amap_tail, x.dt, x.p, 4, dims, str, boxt, q->row_bytes)) return false;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)(9 * Cin), (cuuint64_t)Cout};
        cuuint64_t str[1] = {(cuuint64_t)(9 * Cin) * 2};
        cuuint32_t box[2] = {(cuuint32_t)q->kc, (cuuint32_t)(q->N / q->cl)};   // cluster mode: each CTA loads half of a block
        if (!encode_map(&q->wmap, x.dt, W, 2, dims, str, box, q->row_bytes)) return false;
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)T, (cuuint64_t)F, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)out.ld * 2, (cuuint64_t)T * out.ld * 2, (cuuint64_t)F * T * out.ld * 2};
        cuuint32_t box[4] = {(cuuint32_t)q->panel_cols, (cuuint32_t)(q->case_b ? q->P : 128), 1, (cuuint32_t)(q->case_b ? q->nb : 1)};
        if (!encode_map(&q->omap, x.dt, out.p, 4, dims, str, box, q->panel_bytes)) return false;
    }
    q->rmap = q->omap;
    if (res) {
        cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)T, (cuuint64_t)F, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)res->ld * 2, (cuuint64_t)T * res->ld * 2, (cuuint64_t)F * T * res->ld * 2};
        cuuint32_t box[4] = {(cuuint32_t)q->panel_cols, (cuuint32_t)(q->case_b ? q->P : 128), 1, (cuuint32_t)(q->case_b ? q->nb : 1)};
        if (!encode_map(&q->rmap, x.dt, res->p, 4, dims, str, box, q->panel_bytes)) return false;
    }
    q->res = res ? res->p : nullptr;
    q->res_ld = res ? res->ld : 0;
    q->bias = bias;
    q->relu = relu;   // 0 none, 1 ReLU, 2 Hardtanh(0, 20)
    q->lens = lens;
    if (const char* dbg = getenv("WS_C3_DBG")) q->dbg = atoi(dbg);
    if (getenv("WS_C3_PROF")) {   // tuning aid: synchronous launch + per-role wait-cycle summary on stderr
        long long* prof = nullptr;
        if (cudaMalloc((void**)&prof, (size_t)q->grid * 16 * 8) != cudaSuccess) { set_err("conv3x3: prof buffer"); return false; }
        cudaMemset(prof, 0, (size_t)q->grid * 16 * 8);
        q->prof = prof;
        *op = [q](cudaStream_t s) -> const char* {
            const char* m = ws_c3_launch(q.get(), s);
            if (m) return m;
            cudaStreamSynchronize(s);
            std::vector<long long> h((size_t)q->grid * 16);
            cudaMemcpy(h.data(), q->prof, h.size() * 8, cudaMemcpyDeviceToHost);
            static const char* names[12] = {"prod.wait_aempty", "prod.total", "mma.wait_afull", "mma.wait_tempty", "mma.wait_wfull",
                                            "mma.total", "epi.wait_store", "epi.wait_tfull", "epi.res_issue", "epi.body", "epi.total", "steps"};
            fprintf(stderr, "[c3 prof] grid %d R %d n_mt %d nb %d N %d Cin %d w_res %d stg_bufs %d res %d steps %d:", q->grid, q->R, q->n_mt, q->nb,
                    q->N, q->Cin, q->w_resident, q->stg_bufs, q->res != nullptr, q->total_steps);
            for (int k = 0; k < 12; ++k) {
                double acc = 0;
                for (int c = 0; c < q->grid; ++c) acc += (double)h[(size_t)c * 16 + k];
                fprintf(stderr, " %s=%.0f", names[k], acc / q->grid);
            }
            fprintf(stderr, "\n");
            return nullptr;
        };
        return true;
    }
    *op = [q](cudaStream_t s) { return ws_c3_launch(q.get(), s); };
    if (plan_check_mode())
        set_op_trace("{\"kind\":\"conv3x3\",\"es\":2,\"x\":" + view_json(x) + ",\"out\":" + view_json(out) + ",\"W\":" + jp(W) + ",\"bias\":" + jp(bias) +
                     ",\"res\":" + (res ? view_json(*res) : std::string("null")) + "," + jv("relu", relu) + "," + jv("sf", stride_f) + "," + jv("st", stride_t) +
                     ",\"lens\":" + jp(lens) + "}");
    {
        char buf[200];
        snprintf(buf, sizeof buf, "conv3x3 B=%d F=%d T=%d Cin=%d Cout=%d s=%dx%d caseB=%d nb=%d n_mt=%d R=%d res=%d cl=%d", q->B, q->F, q->T,
                 q->Cin, q->Cout, q->sf, q->st, q->case_b, q->nb, q->n_mt, q->R, q->res != nullptr, q->cl);
        set_op_label(buf, 2.0 * q->B * q->F * q->T * 9.0 * q->Cin * q->Cout);
    }
    return true;
}

extern "C" int ws_cam_smem_bytes(int nmt, int nstages);

bool cam_layer_fill(WsCamLayer* L, int dt, const void* W1, const void* Wl, const float* bn1_scale, const float* bn1_shift,
                    const float* bias2, const float* w1c_t, const float* b1c, const float* w2c_t, const float* b2c, int cin,
                    int dil) {
    memset(L, 0, sizeof(*L));
    if (cin % 32 != 0 || cin > 1024 || dil < 1 || dil > 8) { set_err("cam_dense: cin must be a multiple of 32 (<= 1024), dilation <= 8"); return false; }
    {
        cuuint64_t dims[2] = {(cuuint64_t)cin, 128};
        cuuint64_t str[1] = {(cuuint64_t)cin * 2};
        cuuint32_t box[2] = {64, 128};
        if (!encode_map(&L->w1map, dt, W1, 2, dims, str, box, 128)) return false;
    }
    {
        cuuint64_t dims[2] = {384, 32};
        cuuint64_t str[1] = {384 * 2};
        cuuint32_t box[2] = {64, 32};
        if (!encode_map(&L->wlmap, dt, Wl, 2, dims, str, box, 128)) return false;
    }
    L->bn1_scale = bn1_scale; L->bn1_shift = bn1_shift; L->bias2 = bias2;
    L->w1c_t = w1c_t; L->b1c = b1c; L->w2c_t = w2c_t; L->b2c = b2c;
    L->cin = cin; L->dil = dil;
    return true;
}

bool make_cam_dense_op(const View& X, const WsCamLayer* layers_dev, int l0, int l1, Op* op, bool* unsupported,
                       const int* lens) {
    *unsupported = false;
    const int T = X.T, B = X.B;
    if (X.dt == WS_F32 || X.F != 1 || T < 1 || T > 512 || (X.ld * 2) % 16 != 0 || getenv("WS_NO_CAM_FUSED")) {
        *unsupported = true;
        return false;
    }
    auto q = std::make_shared<WsCamParams>();
    memset(q.get(), 0, sizeof(WsCamParams));
    q->layers = layers_dev; q->l0 = l0; q->l1 = l1; q->lens = lens;
    q->B = B; q->T = T; q->nmt = (T + 127) / 128; q->seg_len = 100; q->dtype = X.dt;
    q->hrows = 16 + 128 * q->nmt;
    int nst = 6;
    while (nst >= 2 && ws_cam_smem_bytes(q->nmt, nst) > ws_cam_max_smem()) --nst;
    if (nst < 2) { *unsupported = true; return false; }
    q->nstages = nst;
    q->smem_bytes = ws_cam_smem_bytes(q->nmt, nst);
    const uint32_t fmt = X.dt == WS_BF16 ? 1u : 0u;
    q->idesc1 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    q->idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    q->grid = B;
    cuuint64_t dims[3] = {(cuuint64_t)X.C, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)X.ld * 2, (cuuint64_t)T * X.ld * 2};
    cuuint32_t boxx[3] = {64, 128, 1}, boxo[3] = {32, 128, 1};
    if (!encode_map(&q->xmap, X.dt, X.p, 3, dims, str, boxx, 128)) return false;
    if (!encode_map(&q->omap, X.dt, X.p, 3, dims, str, boxo, 64)) return false;
    if (getenv("WS_CAM_PROF")) {   // tuning aid: synchronous launch + phase timestamps (cycles since kernel start) of CTA 0
        long long* prof = nullptr;
        if (cudaMalloc((void**)&prof, 16 * 8) != cudaSuccess) { set_err("cam_dense: prof buffer"); return false; }
        cudaMemset(prof, 0, 16 * 8);
        q->prof = prof;
        *op = [q](cudaStream_t s) -> const char* {
            const char* m = ws_cam_launch(q.get(), s);
            if (m) return m;
            cudaStreamSynchronize(s);
            long long h[16];
            cudaMemcpy(h, q->prof, sizeof h, cudaMemcpyDeviceToHost);
            static const char* names[12] = {"params", "transform_done", "acc1", "epi1", "colsums", "ctx", "hid_partial", "hid", "gate",
                                            "acc2", "stored", "complete"};
            fprintf(stderr, "[cam prof] B %d T %d l0 %d nst %d:", q->B, q->T, q->l0, q->nstages);
            for (int k = 0; k < 12; ++k) fprintf(stderr, " %s=%lld", names[k], h[k]);
            fprintf(stderr, "\n");
            return nullptr;
        };
        return true;
    }
    *op = [q](cudaStream_t s) { return ws_cam_launch(q.get(), s); };
    {
        char buf[160];
        snprintf(buf, sizeof buf, "cam_dense layers %d..%d B=%d T=%d", l0, l1 - 1, X.B, X.T);
        set_op_label(buf, 0.0);
    }
    return true;
}

}  // namespace ws

// ------------------------------------------------------------------------------------------------ C ABI: ws_conv
extern "C" int ws_conv(const ws_conv_desc* d, void* stream) {
    using namespace ws;
    if (d == nullptr) { set_err("ws_conv: null descriptor"); return 1; }
    View x;
    x.p = const_cast<void*>(d->x); x.plo = const_cast<void*>(d->x_lo); x.B = d->B; x.F = d->F; x.T = d->T; x.C = d->Cin; x.ld = d->x_ld; x.dt = d->dtype;
    ConvSpec s;
    s.dt = d->dtype;
    int Fo = 0, To = 0;
    const int K = add_conv_taps(s, x, d->kf, d->kt, d->dil_f, d->dil_t, d->pad_f, d->pad_t, d->stride_f, d->stride_t, 0,
                                &Fo, &To);
    if (K < 0) { set_err("ws_conv: too many source planes"); return 1; }
    s.W = d->w; s.Ktot = K; s.Cout = d->Cout;
    s.W_lo = d->w_lo; s.split = (d->x_lo != nullptr);
    s.B = d->B; s.F = Fo; s.T = To;
    s.dense_pointwise = (d->kf == 1 && d->kt == 1 && d->stride_f == 1 && d->stride_t == 1 && d->pad_f == 0 &&
                         d->pad_t == 0);
    View o;
    o.p = d->out; o.plo = d->out_lo; o.B = d->B; o.F = Fo; o.T = To; o.C = d->Cout; o.ld = d->out_ld; o.dt = d->dtype;
    fill_epi_out(s.epi, o);
    s.epi.bias = d->bias; s.epi.act1 = d->act1; s.epi.scale = d->scale; s.epi.shift = d->shift;
    s.epi.res = d->res; s.epi.res_ld = d->res_ld; s.epi.act2 = d->act2;
    s.epi.colsum = d->colsum; s.epi.colsum_T = d->colsum ? To : 0;
    Op op;
    if (d->use_tc) { WS_CKS(ws_tc_init()); WS_CKS(ws_tc2_init()); WS_CKS(ws_tc3_init()); WS_CKS(ws_c3_init()); WS_CKS(ws_astp_init()); }
    bool done = false;
    if (d->use_tc >= 4 && d->kf == 3 && d->kt == 3 && d->dil_f == 1 && d->dil_t == 1 && d->pad_f == 1 && d->pad_t == 1 &&
        d->stride_f >= 1 && d->stride_f <= 2 && d->stride_t >= 1 && d->stride_t <= 2 && d->scale == nullptr && d->x_lo == nullptr && d->colsum == nullptr &&
        (d->act1 == 0 || d->res == nullptr) && (d->act1 <= 1 || d->act1 == WS_ACT_RELU20) && (d->act2 <= 1 || d->act2 == WS_ACT_RELU20) &&
        (d->act1 == 0 || d->act2 == 0 || d->act1 == d->act2)) {
        // halo-resident 3x3 kernel: act(conv + bias [+ res]); with a residual the activation is act2, else act1 (or act2)
        View r = o;
        r.p = const_cast<void*>(d->res); r.ld = d->res_ld;
        bool unsupported = false;
        const int a = d->act1 | d->act2;   // one of them is set, or both to the same (idempotent) activation
        if (make_conv3x3_op(x, o, d->w, d->bias, d->res ? &r : nullptr, a == WS_ACT_RELU20 ? 2 : (a != 0 ? 1 : 0), &op, &unsupported, d->stride_f,
                            d->stride_t)) done = true;
        else if (!unsupported) return 1;
    }
    if (!done && !make_conv_op(s, d->use_tc >= 4 ? 3 : d->use_tc, &op)) return 1;
    const char* m = op((cudaStream_t)stream);
    if (m != nullptr) { set_err(std::string("ws_conv launch: ") + m); return 1; }
    return 0;
}

extern "C" const char* ws_last_error(void) { return ws::get_err().c_str(); }
extern "C" int ws_version(void) { return 100; }
