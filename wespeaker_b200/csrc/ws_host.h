// Host-side plumbing shared by the engine translation units: error reporting, activation views, the conv
// operator builder (tcgen05/TMA or FFMA), device allocation tracking.
#pragma once
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "ws_kernels.cuh"

namespace ws {

void set_err(const std::string& msg);
// tuning aid: op factories leave a label (+ FLOPs) for the op they just built; Builder::push() attaches it to the plan
void set_op_label(const std::string& name, double flops = 0.0);
bool take_op_label(std::string* name, double* flops);
const std::string& get_err();

#define WS_CK(call)                                                                               \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            ws::set_err(std::string(#call) + ": " + cudaGetErrorString(e__));                     \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)
#define WS_CKS(expr)                                                                              \
    do {                                                                                          \
        const char* m__ = (expr);                                                                 \
        if (m__ != nullptr) {                                                                     \
            ws::set_err(std::string(#expr) + ": " + m__);                                         \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

// Plan-check mode (ws_engine_create_plan_check): launch plans are built without a device - buffers and weights get
// placeholder addresses (never dereferenced), tensor maps are validated against the cuTensorMapEncodeTiled rules instead of
// being encoded, nothing is launched.  Thread-local, set for the duration of a plan build.
bool plan_check_mode();
struct PlanCheckScope {
    bool prev;
    explicit PlanCheckScope(bool on);
    ~PlanCheckScope();
};
void* plan_check_alloc(size_t bytes);   // 256-byte aligned placeholder address
// Plan trace (plan-check mode only, ws_engine_plan_trace): op factories leave a JSON description of the op they just built
// (operands as placeholder addresses, shapes, strides, epilogue) and Builder::push() attaches it to the plan, so that a test
// can re-evaluate the plan's arithmetic on the host and compare it with the oracle.  Outside plan-check mode these are no-ops.
void set_op_trace(const std::string& json);
bool take_op_trace(std::string* json);

// channels-last activation view: element (b,f,t,c) at p[((b*F+f)*T+t)*ld + c]
struct View {
    void* p = nullptr;
    void* plo = nullptr;  // 3xTF32 mode: twin buffer holding x - tf32_trunc(x)
    int B = 0, F = 1, T = 0, C = 0;
    long long ld = 0;
    int dt = WS_F32;
    long long npos() const { return (long long)B * F * T; }
    View ch(int c0, int c) const {
        View v = *this;
        v.p = (char*)p + (size_t)c0 * ws_esize(dt);
        if (plo) v.plo = (char*)plo + (size_t)c0 * ws_esize(dt);
        v.C = c;
        return v;
    }
};

// one fused conv/GEMM launch
struct ConvSpec {
    WsSrc src[WS_MAX_SRC];
    int nsrc = 0;
    std::vector<WsTap> taps;
    const void* W = nullptr;  // [Cout][Ktot] in activation dtype
    const void* W_lo = nullptr;  // 3xTF32: W - tf32_trunc(W)
    bool split = false;          // 3xTF32 error-compensated GEMM (fp32 activations with lo twins)
    int Ktot = 0, Cout = 0;
    int B = 0, F = 1, T = 0;  // output extents
    int dt = WS_F32;
    WsEpi epi{};
    bool dense_pointwise = false;  // every tap has dt=df=0 on a dense src with the output's extents -> flatten positions
};

// Append the taps of a (kf x kt) conv over `x` to `spec` (weights columns start at wk0, tap-major then channel).
// Strided convs are expressed with parity-plane source views so every tap is unit-stride (TMA-friendly).
// Returns the number of K columns consumed (kf*kt*x.C) and sets Fo/To.
int add_conv_taps(ConvSpec& spec, const View& x, int kf, int kt, int dil_f, int dil_t, int pad_f, int pad_t,
                  int stride_f, int stride_t, int wk0, int* Fo, int* To);

using Op = std::function<const char*(cudaStream_t)>;
// Build a launch closure for `spec`; use_tc selects the tcgen05 kernel (needs dt-aligned shapes) else FFMA.
// Returns false and sets the error string if the spec cannot be mapped.
bool make_conv_op(const ConvSpec& spec, int use_tc, Op* out);  // 0 FFMA, 1 tcgen05 v1, 2 tcgen05 v2 (falls back to v1)

void fill_epi_out(WsEpi& e, const View& out);

// Fused Res2 chain over a whole SE_Res2Block stage (7 dilated k=3 convs on w8-channel groups, ecapa_tdnn.py:29-78).
// x: block input (B,1,T,8*w8), out: block output buffer (groups 0..6 are written).  W7: [7*w8][3*w8] packed weights in
// the activation dtype; bias/scale/shift: [7][w8] fp32.  Returns false with *unsupported=true when the shape/dtype is
// outside the fused kernel's envelope (16-bit activations, w8 in {64,128}, T <= 256).
bool make_res2_op(const View& x, const View& out, const void* W7, const float* bias, const float* scale,
                  const float* shift, int w8, int dil, Op* op, bool* unsupported, const int* lens = nullptr);

// Fused ASTP tail (ws_astp_fused.cu): x = frame-level features (B,1,T,C), h = tanh(linear1(x)) (B,1,T,128), W2 = linear2 weight
// [C][128] in the activation dtype, stats = fp32 [B][2C] (weighted mean, std).  linear2's bias cancels in the softmax over
// time and is not needed.  *unsupported = true outside the envelope (fp32 activations, C % 128, hidden width != 128).
bool make_astp_op(const View& x, const View& h, const void* W2, float* stats, Op* op, bool* unsupported,
                  const int* lens = nullptr);

// Halo-resident 3x3 pad-1 conv (ws_conv3x3.cu), strides 1 or 2 per axis: out = act(conv(x, W) + bias [+ res]).  x/out/res:
// channels-last 16-bit views (out / res with the strided extents); W: [Cout][9*Cin] tap-major in the activation dtype; relu: 0 none,
// 1 ReLU, 2 Hardtanh(0, 20).  Returns false with
// *unsupported = true when the shape is outside the kernel's envelope (the caller then uses make_conv_op).
bool make_conv3x3_op(const View& x, const View& out, const void* W, const float* bias, const View* res, int relu,
                     Op* op, bool* unsupported, int stride_f = 1, int stride_t = 1, const int* lens = nullptr);

// Fused CAM++ dense layers (ws_cam_dense.cu).  cam_layer_fill builds one host-side layer descriptor (weights are device
// pointers in the activation dtype: W1 [128][cin] with BN2 folded, Wl [32][3*128] tap-major; everything else fp32 device
// arrays).  make_cam_dense_op launches layers [l0, l1) of a device-resident descriptor array over the concat buffer X
// (B,1,T,Cmax).  *unsupported = true when the shape is outside the kernel's envelope (T > 512, fp32 activations).
bool cam_layer_fill(WsCamLayer* L, int dt, const void* W1, const void* Wl, const float* bn1_scale, const float* bn1_shift,
                    const float* bias2, const float* w1c_t, const float* b1c, const float* w2c_t, const float* b2c, int cin,
                    int dil);
bool make_cam_dense_op(const View& X, const WsCamLayer* layers_dev, int l0, int l1, Op* op, bool* unsupported,
                       const int* lens = nullptr);

}  // namespace ws
