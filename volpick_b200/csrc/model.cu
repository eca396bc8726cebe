// Model handle, host-side weight folding / packing, and the forward plans.
//
// vp_model_create   <- SeisBenchModel.from_pretrained -> load_state_dict
//                      (/root/reference/README.md:46-47, /root/reference/volpick/model/train.py:94)
// vp_forward        <- EQTransformer.forward / PhaseNet.forward (SeisBench; SURVEY.md Appendix A / B;
//                      call sites /root/reference/volpick/model/eval_taks0.py:68-72,85-89)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "fused.cuh"
#include "tcconv.cuh"

namespace vp {

// ---- error / launch accounting ------------------------------------------------------------------
static thread_local std::string g_error;
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}
void count_launch(int n) { g_launches += n; }

namespace {
std::mutex g_attr_mutex;
std::map<std::pair<const void *, int>, size_t> g_attr_granted;
int g_sm_count[64];  // 0: not queried yet
}  // namespace

int ensure_dyn_smem(const void *kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return VP_OK;
    int dev = 0;
    VP_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    size_t &granted = g_attr_granted[std::make_pair(kernel, dev)];
    if (bytes > granted) {
        VP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        granted = bytes;
    }
    return VP_OK;
}

int device_sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    if (g_sm_count[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_sm_count[dev] = n;
    }
    return g_sm_count[dev];
}

thread_local bool g_ktimer_on = false;
namespace {
struct KSpan {
    int cls;
    cudaEvent_t e0, e1;
};
thread_local std::vector<KSpan> g_kspans;
thread_local std::vector<cudaEvent_t> g_kevent_pool;
cudaEvent_t ktimer_event() {
    if (!g_kevent_pool.empty()) {
        cudaEvent_t e = g_kevent_pool.back();
        g_kevent_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
const char *const KCLASS_NAMES =
    "slice_normalize,slice_enc0,tcconv,deca,decb,conv1d_f32,convt_f32,lstm,attention,pack_cl16,stack,nan_bounds,pick,sosfilt,resstack";
}  // namespace
void ktimer_mark(int cls, cudaStream_t s, bool end) {
    if (!end) {
        KSpan sp{cls, ktimer_event(), nullptr};
        cudaEventRecord(sp.e0, s);
        g_kspans.push_back(sp);
    } else {
        for (size_t i = g_kspans.size(); i-- > 0;) {
            if (g_kspans[i].cls == cls && g_kspans[i].e1 == nullptr) {
                g_kspans[i].e1 = ktimer_event();
                cudaEventRecord(g_kspans[i].e1, s);
                break;
            }
        }
    }
}

constexpr double BN_EPS = 1e-3;
constexpr int64_t EQT_FLOATS = 378823;
constexpr int64_t PN_FLOATS = 269675;
constexpr int64_t MAX_CHUNK = 4096;  // windows per internal forward chunk (grid.y limit is 65535)

struct Cursor {
    const float *p;
    int64_t left;
    const float *take(int64_t n) {
        const float *r = p;
        p += n;
        left -= n;
        return r;
    }
};

struct BN {
    const float *w, *b, *rm, *rv;
    int c;
};
static BN take_bn(Cursor &cur, int c) {
    BN bn;
    bn.w = cur.take(c);
    bn.b = cur.take(c);
    bn.rm = cur.take(c);
    bn.rv = cur.take(c);
    bn.c = c;
    return bn;
}
// y = x * scale + shift, folded in double (a denormal running_var must not be flushed on the device)
static void bn_affine(const BN &bn, std::vector<double> &scale, std::vector<double> &shift) {
    scale.resize(bn.c);
    shift.resize(bn.c);
    for (int i = 0; i < bn.c; ++i) {
        scale[i] = (double)bn.w[i] / std::sqrt((double)bn.rv[i] + BN_EPS);
        shift[i] = (double)bn.b[i] - (double)bn.rm[i] * scale[i];
    }
}

struct Packed {
    std::vector<float> host;
    int64_t add(int64_t n) {
        const int64_t off = align_up((int64_t)host.size(), 64);  // 256-byte aligned blocks
        host.resize(off + n, 0.f);
        return off;
    }
};

struct ConvW {
    int64_t w, b;  // offsets into the packed buffer
    int cin, cout, coutp, k;
};
// W (cout, cin, k) [+ bias] [+ BatchNorm after the conv] -> wt[cin][k][coutp], b[coutp]
static ConvW pack_conv(Packed &pk, const float *W, const float *bias, const BN *bn, int cout, int cin, int k) {
    ConvW cw;
    cw.cin = cin;
    cw.cout = cout;
    cw.k = k;
    cw.coutp = (cout + 3) / 4 * 4;
    std::vector<double> sc(cout, 1.0), sh(cout, 0.0);
    if (bn) bn_affine(*bn, sc, sh);
    cw.w = pk.add((int64_t)cin * k * cw.coutp);
    cw.b = pk.add(cw.coutp);
    for (int co = 0; co < cout; ++co) {
        for (int ci = 0; ci < cin; ++ci)
            for (int kk = 0; kk < k; ++kk)
                pk.host[cw.w + ((int64_t)ci * k + kk) * cw.coutp + co] =
                    (float)((double)W[((int64_t)co * cin + ci) * k + kk] * sc[co]);
        const double b0 = bias ? (double)bias[co] : 0.0;
        pk.host[cw.b + co] = (float)(b0 * sc[co] + sh[co]);
    }
    return cw;
}
// ConvTranspose1d W (cin, cout, 7) + BN -> wt[cin][7][cout], b[cout]
static ConvW pack_convt(Packed &pk, const float *W, const BN &bn, int cin, int cout) {
    ConvW cw;
    cw.cin = cin;
    cw.cout = cw.coutp = cout;
    cw.k = 7;
    std::vector<double> sc, sh;
    bn_affine(bn, sc, sh);
    cw.w = pk.add((int64_t)cin * 7 * cout);
    cw.b = pk.add(cout);
    for (int ci = 0; ci < cin; ++ci)
        for (int co = 0; co < cout; ++co)
            for (int kk = 0; kk < 7; ++kk)
                pk.host[cw.w + ((int64_t)ci * 7 + kk) * cout + co] =
                    (float)((double)W[((int64_t)ci * cout + co) * 7 + kk] * sc[co]);
    for (int co = 0; co < cout; ++co) pk.host[cw.b + co] = (float)sh[co];
    return cw;
}
struct AffW {
    int64_t scale, shift;
};
static AffW pack_affine(Packed &pk, const BN &bn) {
    std::vector<double> sc, sh;
    bn_affine(bn, sc, sh);
    AffW a;
    a.scale = pk.add(bn.c);
    a.shift = pk.add(bn.c);
    for (int i = 0; i < bn.c; ++i) {
        pk.host[a.scale + i] = (float)sc[i];
        pk.host[a.shift + i] = (float)sh[i];
    }
    return a;
}
struct LstmW {
    int64_t ih, hh, b;
    int cin, ndir;
};
// per direction: w_ih (64, cin), w_hh (64, 16), b_ih (64), b_hh (64); PyTorch gate order i,f,g,o
static void pack_lstm_dir(Packed &pk, const LstmW &lw, int dir, Cursor &cur) {
    const int cin = lw.cin;
    const float *wih = cur.take(64 * cin), *whh = cur.take(64 * 16), *bih = cur.take(64), *bhh = cur.take(64);
    for (int ci = 0; ci < cin; ++ci)
        for (int j = 0; j < 16; ++j)
            for (int g = 0; g < 4; ++g)
                pk.host[lw.ih + (((int64_t)dir * cin + ci) * 16 + j) * 4 + g] = wih[(g * 16 + j) * cin + ci];
    for (int k = 0; k < 16; ++k)
        for (int j = 0; j < 16; ++j)
            for (int g = 0; g < 4; ++g)
                pk.host[lw.hh + (((int64_t)dir * 16 + k) * 16 + j) * 4 + g] = whh[(g * 16 + j) * 16 + k];
    for (int j = 0; j < 16; ++j)
        for (int g = 0; g < 4; ++g)
            pk.host[lw.b + ((int64_t)dir * 16 + j) * 4 + g] = bih[g * 16 + j] + bhh[g * 16 + j];
}
static LstmW alloc_lstm(Packed &pk, int cin, int ndir) {
    LstmW lw;
    lw.cin = cin;
    lw.ndir = ndir;
    lw.ih = pk.add((int64_t)ndir * cin * 64);
    lw.hh = pk.add((int64_t)ndir * 16 * 64);
    lw.b = pk.add((int64_t)ndir * 64);
    return lw;
}
// attention: Wx (16,32), Wt (16,32), bh (32), Wa (32,1), ba (1) -> AW_* block
static void pack_attention(Packed &pk, int64_t off, Cursor &cur) {
    const float *Wx = cur.take(512), *Wt = cur.take(512), *bh = cur.take(32), *Wa = cur.take(32), *ba = cur.take(1);
    std::memcpy(&pk.host[off + AW_WT], Wt, 512 * sizeof(float));
    std::memcpy(&pk.host[off + AW_WX], Wx, 512 * sizeof(float));
    std::memcpy(&pk.host[off + AW_BH], bh, 32 * sizeof(float));
    std::memcpy(&pk.host[off + AW_WA], Wa, 32 * sizeof(float));
    pk.host[off + AW_BA] = ba[0];
}

}  // namespace vp

using namespace vp;

struct vp_model {
    int kind = 0;
    int device = 0;
    int in_samples = 0;
    float *d_weights = nullptr;
    int64_t n_packed = 0;
    // EQTransformer
    ConvW enc[7];
    struct {
        AffW n1, n2;
        ConvW c1, c2;
    } res[7];
    struct {
        LstmW lstm;
        ConvW conv;
    } bil[3];
    int64_t tr[2] = {0, 0};  // AW blocks (transformer_d0, transformer_d)
    ConvW dec[3][7];         // [decoder_d, pick_decoders.0, pick_decoders.1][layer]; packed layer-major
    ConvW head[3];
    LstmW pick_lstm[2];
    int64_t pick_attn[2] = {0, 0};
    // PhaseNet
    ConvW inc, down_same[5], down_down[4], up_t[4], up_same[4], outc;
    std::string tap_names;
    float enc0_w[8 * 3 * 11] = {0}, enc0_b[8] = {0};  // host copy of encoder.convs.0 for the fused slicer kernel (kernel parameter)
    // tensor-core (tcgen05) weight sets: [0] = fp16 hi/lo split (f16x3), [1] = bf16
    struct TcSet {
        TcLayer enc[7], dec[7], head;
        TcLayer encf[3];           // encoder.convs.1 / .2 with 4 time steps folded into the MMA K / N (index = layer; see tc_fold4_same)
        TcLayer res1[7], res2[7];  // res-CNN convs (BatchNorm + ReLU of their inputs live in the producer's epilogue)
        TcLayer lproj;             // bi_lstm_stack.members.0: input projection of both directions as a 1x1 conv (64 -> 2 x 64 gates)
        uint16_t *d_w = nullptr;
        float *d_b = nullptr;
        std::vector<float> res_par;   // epilogue parameters of the on-chip res-CNN stack (fused_res2.cu), host copy
        float *d_res_par = nullptr;
        bool ready = false;
        DecBPlan decb;  // fused decoder tail (fused_dec.cu)
        DecB2Plan decb2;  // fused decoder tail with the wide levels in tensor memory (fused_dec2.cu): the default
        EncAPlan enca;    // fused encoder front, convs.1-3 (fused_enc.cu)
        DecAPlan deca;  // fused decoder middle, convs.1 + convs.2 (fused_deca.cu)
    } tc[2];
    // PhaseNet on the tensor cores (same precision sets).  Stride-4 convs and the stride-4 ConvTranspose1d run as k = 2
    // convs on the row-reshaped channel-last buffers ([T][C] seen as [T / 4][4 C]); see build_pn_tc.
    struct PnTcSet {
        TcLayer inc, ds[5], dd[4], ut[4], us[4];
        TcLayer ds0f, us3f;  // level-0 'same' convs with 4 time steps folded into the MMA K / N (see fold_same_conv)
        uint16_t *d_w = nullptr;
        float *d_b = nullptr;
        bool ready = false;
    } pn_tc[2];
    float pn_head_w[24] = {0}, pn_head_b[3] = {0};  // `out` conv (3, 8, 1) + bias, fused into the last layer's epilogue
    float pn_inc_w[8 * 3 * 7] = {0}, pn_inc_b[8] = {0};  // `inc` conv + in_bn folded, for the fused slicer kernel
};

template <class Set>
static int upload_tc_layers(Set &ts, const std::vector<TcLayer *> &layers) {
    size_t nw = 0, nb = 0;
    for (TcLayer *L : layers) {
        L->w_off = (int64_t)nw;
        L->b_off = (int64_t)nb;
        nw += (L->blocks.size() + 63) / 64 * 64;
        nb += (L->bias.size() + 63) / 64 * 64;
    }
    VP_CUDA_CHECK(cudaMalloc(&ts.d_w, nw * sizeof(uint16_t) + 256));
    VP_CUDA_CHECK(cudaMalloc(&ts.d_b, nb * sizeof(float) + 256));
    for (TcLayer *L : layers) {
        VP_CUDA_CHECK(cudaMemcpy(ts.d_w + L->w_off, L->blocks.data(), L->blocks.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        VP_CUDA_CHECK(cudaMemcpy(ts.d_b + L->b_off, L->bias.data(), L->bias.size() * sizeof(float), cudaMemcpyHostToDevice));
        L->blocks.clear();
        L->blocks.shrink_to_fit();
    }
    ts.ready = true;
    return VP_OK;
}
static int upload_pn_tc(vp_model::PnTcSet &ts) {
    std::vector<TcLayer *> layers{&ts.inc, &ts.ds0f, &ts.us3f};
    for (int i = 0; i < 5; ++i) layers.push_back(&ts.ds[i]);
    for (int i = 0; i < 4; ++i) {
        layers.push_back(&ts.dd[i]);
        layers.push_back(&ts.ut[i]);
        layers.push_back(&ts.us[i]);
    }
    return upload_tc_layers(ts, layers);
}

static int upload_tc(vp_model::TcSet &ts) {
    std::vector<TcLayer *> layers;
    for (int i = 0; i < 7; ++i) layers.push_back(&ts.enc[i]);
    for (int i = 1; i < 3; ++i) layers.push_back(&ts.encf[i]);
    for (int i = 0; i < 7; ++i) layers.push_back(&ts.dec[i]);
    layers.push_back(&ts.head);
    layers.push_back(&ts.lproj);
    for (int i = 0; i < 7; ++i) {
        layers.push_back(&ts.res1[i]);
        layers.push_back(&ts.res2[i]);
    }
    size_t nw = 0, nb = 0;
    for (TcLayer *L : layers) {
        L->w_off = (int64_t)nw;
        L->b_off = (int64_t)nb;
        nw += (L->blocks.size() + 63) / 64 * 64;
        nb += (L->bias.size() + 63) / 64 * 64;
    }
    VP_CUDA_CHECK(cudaMalloc(&ts.d_w, nw * sizeof(uint16_t) + 256));
    VP_CUDA_CHECK(cudaMalloc(&ts.d_b, nb * sizeof(float) + 256));
    for (TcLayer *L : layers) {
        VP_CUDA_CHECK(cudaMemcpy(ts.d_w + L->w_off, L->blocks.data(), L->blocks.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        VP_CUDA_CHECK(cudaMemcpy(ts.d_b + L->b_off, L->bias.data(), L->bias.size() * sizeof(float), cudaMemcpyHostToDevice));
        L->blocks.clear();
        L->blocks.shrink_to_fit();
    }
    VP_CUDA_CHECK(cudaMalloc(&ts.d_res_par, ts.res_par.size() * sizeof(float) + 256));
    VP_CUDA_CHECK(cudaMemcpy(ts.d_res_par, ts.res_par.data(), ts.res_par.size() * sizeof(float), cudaMemcpyHostToDevice));
    ts.ready = true;
    return VP_OK;
}

static const int kEncC[8] = {3, 8, 16, 16, 32, 32, 64, 64};
static const int kEncK[7] = {11, 9, 7, 7, 5, 5, 3};
static const int kResK[7] = {3, 3, 3, 3, 2, 3, 2};
static const int kDecC[8] = {16, 64, 64, 32, 32, 16, 16, 8};
static const int kDecK[7] = {3, 5, 5, 7, 7, 9, 11};
static const int kPnC[5] = {8, 16, 32, 64, 128};

static int build_eqt(vp_model *m, Cursor &cur, Packed &pk) {
    const float *eW[7], *eB[7];
    for (int i = 0; i < 7; ++i) {
        const float *W = cur.take((int64_t)kEncC[i + 1] * kEncC[i] * kEncK[i]);
        const float *b = cur.take(kEncC[i + 1]);
        eW[i] = W;
        eB[i] = b;
        if (i == 0) {
            std::memcpy(m->enc0_w, W, sizeof(m->enc0_w));
            std::memcpy(m->enc0_b, b, sizeof(m->enc0_b));
        }
        m->enc[i] = pack_conv(pk, W, b, nullptr, kEncC[i + 1], kEncC[i], kEncK[i]);
    }
    const float *rW[7][2], *rB[7][2];
    for (int i = 0; i < 7; ++i) {
        BN n1 = take_bn(cur, 64);
        const float *W1 = cur.take(64 * 64 * kResK[i]), *b1 = cur.take(64);
        BN n2 = take_bn(cur, 64);
        const float *W2 = cur.take(64 * 64 * kResK[i]), *b2 = cur.take(64);
        rW[i][0] = W1, rB[i][0] = b1, rW[i][1] = W2, rB[i][1] = b2;
        m->res[i].n1 = pack_affine(pk, n1);
        m->res[i].c1 = pack_conv(pk, W1, b1, nullptr, 64, 64, kResK[i]);
        m->res[i].n2 = pack_affine(pk, n2);
        m->res[i].c2 = pack_conv(pk, W2, b2, nullptr, 64, 64, kResK[i]);
    }
    for (int i = 0; i < 3; ++i) {
        const int cin = i == 0 ? 64 : 16;
        m->bil[i].lstm = alloc_lstm(pk, cin, 2);
        pack_lstm_dir(pk, m->bil[i].lstm, 0, cur);
        pack_lstm_dir(pk, m->bil[i].lstm, 1, cur);
        const float *W = cur.take(16 * 32), *b = cur.take(16);
        BN bn = take_bn(cur, 16);
        m->bil[i].conv = pack_conv(pk, W, b, &bn, 16, 32, 1);
    }
    for (int i = 0; i < 2; ++i) {
        const int64_t off = pk.add(AW_SIZE);
        m->tr[i] = off;
        pack_attention(pk, off, cur);
        const float *g1 = cur.take(16), *b1 = cur.take(16);
        const float *l1w = cur.take(128 * 16), *l1b = cur.take(128);
        const float *l2w = cur.take(16 * 128), *l2b = cur.take(16);
        const float *g2 = cur.take(16), *b2 = cur.take(16);
        std::memcpy(&pk.host[off + AW_G1], g1, 16 * sizeof(float));
        std::memcpy(&pk.host[off + AW_B1], b1, 16 * sizeof(float));
        std::memcpy(&pk.host[off + AW_L1W], l1w, 2048 * sizeof(float));
        std::memcpy(&pk.host[off + AW_L1B], l1b, 128 * sizeof(float));
        for (int c = 0; c < 16; ++c)
            for (int mm = 0; mm < 128; ++mm) pk.host[off + AW_L2W + mm * 16 + c] = l2w[c * 128 + mm];
        std::memcpy(&pk.host[off + AW_L2B], l2b, 16 * sizeof(float));
        std::memcpy(&pk.host[off + AW_G2], g2, 16 * sizeof(float));
        std::memcpy(&pk.host[off + AW_B2], b2, 16 * sizeof(float));
    }
    // state-dict order: decoder_d.*, conv_d, pick_lstms.*, pick_attentions.*, pick_decoders.*, pick_convs.*
    const float *dW[3][7], *dB[3][7], *hW[3], *hB[3];
    for (int i = 0; i < 7; ++i) {
        dW[0][i] = cur.take((int64_t)kDecC[i + 1] * kDecC[i] * kDecK[i]);
        dB[0][i] = cur.take(kDecC[i + 1]);
    }
    hW[0] = cur.take(88);
    hB[0] = cur.take(1);
    for (int g = 0; g < 2; ++g) {
        m->pick_lstm[g] = alloc_lstm(pk, 16, 1);
        pack_lstm_dir(pk, m->pick_lstm[g], 0, cur);
    }
    for (int g = 0; g < 2; ++g) {
        m->pick_attn[g] = pk.add(AW_SIZE);
        pack_attention(pk, m->pick_attn[g], cur);
    }
    for (int g = 1; g < 3; ++g)
        for (int i = 0; i < 7; ++i) {
            dW[g][i] = cur.take((int64_t)kDecC[i + 1] * kDecC[i] * kDecK[i]);
            dB[g][i] = cur.take(kDecC[i + 1]);
        }
    for (int g = 1; g < 3; ++g) {
        hW[g] = cur.take(88);
        hB[g] = cur.take(1);
    }
    // layer-major packing: the three decoders of one layer are equally sized consecutive blocks,
    // so one grouped launch addresses them with a constant group stride
    for (int i = 0; i < 7; ++i)
        for (int g = 0; g < 3; ++g)
            m->dec[g][i] = pack_conv(pk, dW[g][i], dB[g][i], nullptr, kDecC[i + 1], kDecC[i], kDecK[i]);
    for (int g = 0; g < 3; ++g) m->head[g] = pack_conv(pk, hW[g], hB[g], nullptr, 1, 8, 11);
    m->tap_names =
        "enc0,enc1,enc2,enc3,enc4,enc5,enc6,res0,res1,res2,res3,res4,res5,res6,bilstm0_lstm,bilstm0,"
        "bilstm1_lstm,bilstm1,bilstm2_lstm,bilstm2,transformer_d0,transformer_d,pick_lstm,pick_attn,"
        "dec0,dec1,dec2,dec3,dec4,dec5,dec6";
    // ---- tensor-core weight sets (tcconv.cu): encoder direct convs, decoders with the x2 up-sampling
    // folded into the weights (layer 2 keeps the explicit up-sampling because of its crop), heads
    for (int set = 0; set < 2; ++set) {
        const int split = set == 0 ? 2 : 1;
        vp_model::TcSet &ts = m->tc[set];
        for (int i = 0; i < 7; ++i) {
            int cin = kEncC[i];
            std::vector<float> wpad;
            const float *W = eW[i];
            if (cin == 3) {  // pad the 3 input components to one 8-channel plane
                wpad.assign((size_t)kEncC[i + 1] * 8 * kEncK[i], 0.f);
                for (int co = 0; co < kEncC[i + 1]; ++co)
                    for (int ci = 0; ci < 3; ++ci)
                        for (int kk = 0; kk < kEncK[i]; ++kk)
                            wpad[((size_t)co * 8 + ci) * kEncK[i] + kk] = W[((size_t)co * 3 + ci) * kEncK[i] + kk];
                W = wpad.data();
                cin = 8;
            }
            const float *wl[1] = {W}, *bl[1] = {eB[i]};
            int rc = tc_build_layer(ts.enc[i], TC_DIRECT, cin, kEncC[i + 1], kEncK[i], 0, split, 1, wl, bl);
            if (rc != VP_OK) return rc;
            if (i == 1 || i == 2) {  // 8 -> 16 (k9) and 16 -> 16 (k7): folded x4 -> 32 / 64 -> 64 columns, 3 row taps
                std::vector<float> wf, bf;
                tc_fold4_same(W, eB[i], kEncC[i + 1], cin, kEncK[i], wf, bf);
                const float *wfl[1] = {wf.data()}, *bfl[1] = {bf.data()};
                rc = tc_build_layer(ts.encf[i], TC_DIRECT, 4 * cin, 4 * kEncC[i + 1], 3, 0, split, 1, wfl, bfl, 1);
                if (rc != VP_OK) return rc;
            }
        }
        {   // fused encoder front: convs.1 from the raw (16, 8, 9) weights, convs.2 / .3 from their tensor-core layers
            int rc = enca_build(ts.enca, ts.enc[2], ts.enc[3], eW[1], eB[1], split);
            if (rc != VP_OK) return rc;
        }
        for (int i = 0; i < 7; ++i) {
            const float *wl[3] = {dW[0][i], dW[1][i], dW[2][i]}, *bl[3] = {dB[0][i], dB[1][i], dB[2][i]};
            const int mode = (i == 2) ? TC_DIRECT_UPS : TC_POLYPHASE;  // decoder stage 2 crops one sample (376 -> 375)
            int rc = tc_build_layer(ts.dec[i], mode, kDecC[i], kDecC[i + 1], kDecK[i], i == 2 ? 1 : 0, split, 3, wl, bl);
            if (rc != VP_OK) return rc;
        }
        const float *wl[3] = {hW[0], hW[1], hW[2]}, *bl[3] = {hB[0], hB[1], hB[2]};
        int rc = tc_build_layer(ts.head, TC_DIRECT, 8, 1, 11, 0, split, 3, wl, bl);
        if (rc != VP_OK) return rc;
        for (int i = 0; i < 7; ++i)
            for (int j = 0; j < 2; ++j) {  // k = 3: pad (1, 1); k = 2: SeisBench pads one zero on the right only
                const float *w1[1] = {rW[i][j]}, *b1[1] = {rB[i][j]};
                rc = tc_build_layer(j == 0 ? ts.res1[i] : ts.res2[i], TC_DIRECT, 64, 64, kResK[i], 0, split, 1, w1, b1, kResK[i] == 3 ? 1 : 0);
                if (rc != VP_OK) return rc;
            }
        {   // epilogue parameters of the on-chip stack: conv biases, folded pre-activation BatchNorms, cumulative conv2 biases
            const float *b1[7], *b2[7], *n1s[7], *n1h[7], *n2s[7], *n2h[7];
            for (int i = 0; i < 7; ++i) {
                b1[i] = ts.res1[i].bias.data();
                b2[i] = ts.res2[i].bias.data();
                n1s[i] = pk.host.data() + m->res[i].n1.scale;
                n1h[i] = pk.host.data() + m->res[i].n1.shift;
                n2s[i] = pk.host.data() + m->res[i].n2.scale;
                n2h[i] = pk.host.data() + m->res[i].n2.shift;
            }
            resstack2_params(b1, b2, n1s, n1h, n2s, n2h, ts.res_par);
        }
        {   // BiLSTM block 0 input projection: column n = dir * 64 + unit * 4 + gate (the order lstm_kernel reads), bias b_ih + b_hh
            const LstmW &lw = m->bil[0].lstm;
            std::vector<float> wp((size_t)128 * 64), bp(128);
            for (int dir = 0; dir < 2; ++dir)
                for (int j = 0; j < 16; ++j)
                    for (int g = 0; g < 4; ++g) {
                        const int n = dir * 64 + j * 4 + g;
                        bp[n] = pk.host[lw.b + ((int64_t)dir * 16 + j) * 4 + g];
                        for (int ci = 0; ci < 64; ++ci) wp[(size_t)n * 64 + ci] = pk.host[lw.ih + (((int64_t)dir * 64 + ci) * 16 + j) * 4 + g];
                    }
            const float *w1[1] = {wp.data()}, *b1[1] = {bp.data()};
            rc = tc_build_layer(ts.lproj, TC_DIRECT, 64, 128, 1, 0, split, 1, w1, b1);
            if (rc != VP_OK) return rc;
        }
        // fused decoder tail: (1, 8, 11) head weights -> [c * 11 + k]; tile = 47 / 75 rows of the 375-sample level
        float head_w[3][88], head_b[3];
        for (int g = 0; g < 3; ++g) {
            std::memcpy(head_w[g], hW[g], 88 * sizeof(float));
            head_b[g] = hB[g][0];
        }
        // rows of the 375-sample level per work item: 6 (f16x3) / 4 (bf16) items cover the 5000 kept samples of a blinded window
        int tile_m = split == 2 ? 53 : 79;
        if (const char *e = getenv(split == 2 ? "VP_DECB_M2" : "VP_DECB_M1")) tile_m = atoi(e);  // tuning / debugging aid
        rc = decb_build(ts.decb, ts.dec, split, tile_m, head_w, head_b);
        if (rc != VP_OK) return rc;
        {
            const float *w4[3] = {dW[0][4], dW[1][4], dW[2][4]}, *b4[3] = {dB[0][4], dB[1][4], dB[2][4]};
            rc = decb2_build(ts.decb2, ts.dec, w4, b4, split, head_w, head_b);
            if (rc != VP_OK) return rc;
        }
        {   // fused decoder middle: decoder.convs.2 in polyphase form (the crop is corrected inside the kernel)
            TcLayer dec2p;
            const float *w2[3] = {dW[0][2], dW[1][2], dW[2][2]}, *b2[3] = {dB[0][2], dB[1][2], dB[2][2]};
            rc = tc_build_layer(dec2p, TC_POLYPHASE, kDecC[2], kDecC[3], kDecK[2], 0, split, 3, w2, b2);
            if (rc == VP_OK) rc = deca_build(ts.deca, ts.dec[1], dec2p, split, w2);
            if (rc != VP_OK) return rc;
        }
    }
    return VP_OK;
}

// Conv (cout, cin, k) [+ bias] + BatchNorm folded in double -> fp32 (cout, cin_pad, k) + (cout)
struct FoldedConv {
    std::vector<float> w, b;
};
static FoldedConv fold_conv(const float *W, const float *bias, const BN &bn, int cout, int cin, int k, int cin_pad) {
    FoldedConv f;
    std::vector<double> sc, sh;
    bn_affine(bn, sc, sh);
    f.w.assign((size_t)cout * cin_pad * k, 0.f);
    f.b.assign(cout, 0.f);
    for (int co = 0; co < cout; ++co) {
        for (int ci = 0; ci < cin; ++ci)
            for (int kk = 0; kk < k; ++kk)
                f.w[((size_t)co * cin_pad + ci) * k + kk] = (float)((double)W[((size_t)co * cin + ci) * k + kk] * sc[co]);
        f.b[co] = (float)((bias ? (double)bias[co] : 0.0) * sc[co] + sh[co]);
    }
    return f;
}

// 'same' Conv1d (k = 7, pad 3) with F time steps folded into the channels.  Small-channel layers are bound by the
// shared-memory reads of the A operand (4 KB per tcgen05.mma whatever its N: ncu shows the tc pipe 91 % busy with the
// tensor math at 18 % for N = 16), so fewer, wider MMAs win: on the [T / F][F C] view of the channel-last buffers
// (the same bytes) the conv needs ceil-wise 3 row taps instead of 7 and yields F * C_out columns.
// Source s holds sample t at position t + o_in[s] of its buffer (row (t + o_in) / F, lane (t + o_in) % F), the output
// sample t goes to position t + o_out.  Returns the (F * cout, sum_s F * c_s, 3) weights (+ replicated bias) of the
// row conv and its left padding in rows; the offsets must give every source the same first row tap.
constexpr int PN_FOLD = 4;
static int fold_same_conv(const FoldedConv &f, int cout, const int *c_src, const int *o_in, int n_src, int o_out, FoldedConv &out,
                          int &pad_rows) {
    const int F = PN_FOLD, K = 7, P = 3;
    int cin = 0;
    for (int s = 0; s < n_src; ++s) cin += c_src[s];
    auto fdiv = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
    int d_min = 0, d_max = 0;
    for (int s = 0; s < n_src; ++s) {
        const int a = o_in[s] - P - o_out;  // buffer position (relative to F * row) of the first sample read by lane 0
        const int lo = fdiv(a, F), hi = fdiv(a + (F - 1) + (K - 1), F);
        if (s == 0) d_min = lo, d_max = hi;
        VP_REQUIRE(lo == d_min && hi - lo <= 2, VP_ERR_UNSUPPORTED, "fold_same_conv: source %d needs row taps [%d, %d] (first source: from %d)", s, lo, hi, d_min);
        d_max = std::max(d_max, hi);
    }
    const int ntap = 3;
    VP_REQUIRE(d_max - d_min + 1 <= ntap, VP_ERR_UNSUPPORTED, "fold_same_conv: %d row taps", d_max - d_min + 1);
    pad_rows = -d_min;
    const int cinF = F * cin, coutF = F * cout;
    out.w.assign((size_t)coutF * cinF * ntap, 0.f);
    out.b.assign(coutF, 0.f);
    for (int q = 0; q < F; ++q)
        for (int co = 0; co < cout; ++co) {
            out.b[(size_t)q * cout + co] = f.b[co];
            int ch0 = 0, col0 = 0;  // first channel of the source in the original / folded channel order
            for (int s = 0; s < n_src; ++s) {
                for (int j = 0; j < K; ++j) {
                    const int pos = q + j - P + o_in[s] - o_out;  // source buffer position relative to F * row
                    const int d = fdiv(pos, F), lane = pos - d * F;
                    for (int ci = 0; ci < c_src[s]; ++ci)
                        out.w[(((size_t)q * cout + co) * cinF + col0 + lane * c_src[s] + ci) * ntap + (d - d_min)] =
                            f.w[((size_t)co * cin + ch0 + ci) * K + j];
                }
                ch0 += c_src[s];
                col0 += F * c_src[s];
            }
        }
    return VP_OK;
}

// PhaseNet layers as tensor-core convs.  With channel-last rows a buffer [T][C] is also [T / 4][4 C], so
//   * the stride-4 Conv1d (k = 7, SeisBench's manual left pad pl): out[s] = sum_j w[j] xp[4 s + j] over the padded signal
//     xp becomes a k = 2 'valid' conv over rows of 4 C channels (tap 0: w[0..3], tap 1: w[4..6] and a zero), and
//   * the stride-4 ConvTranspose1d (k = 7): out[4 r + q] = w[q] x[r] + w[q + 4] x[r - 1] becomes a k = 2 conv (left pad 1)
//     with 4 C_out output columns per row, which is again the channel-last layout of the up-sampled signal.
// Wide layers are split along the output columns into `groups` launches-in-one (grid.y) so that the resident weights
// fit in shared memory.
static int build_pn_tc(vp_model *m, const float *incW, const float *incB, const BN &incBN, const float *const *dsW, const BN *dsBN,
                       const float *const *ddW, const BN *ddBN, const float *const *utW, const BN *utBN,
                       const float *const *usW, const BN *usBN) {
    // N groups of the three layers whose weights exceed the shared memory: the stride-4 conv 64 -> 64 (K = 512), the concat conv
    // 128 -> 64 (K = 896) and the transposed conv 128 -> 4 x 64 (K = 256).  Two groups of N = 32 leave room for ONE A stage only,
    // but an N = 16 MMA costs the tensor pipe as much as an N = 32 one and every group re-reads the A tiles: tcconv class 3.24 ->
    // 3.13 ms per station-day for the first two; the transposed conv is faster with four groups of N = 64 (3.27 with two).
    // A/B aid: VP_PN_GROUPS="dd,us,ut".
    int g_dd = 2, g_us = 2, g_ut = 4;
    if (const char *e = getenv("VP_PN_GROUPS")) sscanf(e, "%d,%d,%d", &g_dd, &g_us, &g_ut);
    for (int set = 0; set < 2; ++set) {
        const int split = set == 0 ? 2 : 1;
        vp_model::PnTcSet &ts = m->pn_tc[set];
        auto build = [&](TcLayer &L, const FoldedConv &f, int cin, int cout, int k, int groups, int pad_left) -> int {
            std::vector<const float *> wl(groups), bl(groups);
            const int cg = cout / groups;
            for (int g = 0; g < groups; ++g) {
                wl[g] = f.w.data() + (size_t)g * cg * cin * k;
                bl[g] = f.b.data() + (size_t)g * cg;
            }
            return tc_build_layer(L, TC_DIRECT, cin, cg, k, 0, split, groups, wl.data(), bl.data(), pad_left);
        };
        int rc = build(ts.inc, fold_conv(incW, incB, incBN, 8, 3, 7, 8), 8, 8, 7, 1, 3);
        if (rc != VP_OK) return rc;
        int last = 8;
        for (int i = 0; i < 5; ++i) {
            const int f = kPnC[i];
            rc = build(ts.ds[i], fold_conv(dsW[i], nullptr, dsBN[i], f, last, 7, last), last, f, 7, f == 128 ? 2 : 1, 3);
            if (rc != VP_OK) return rc;
            if (i == 0) {  // folded form: B0 (offset 0) -> skip_0 stored at offset 3 (the stride-4 conv's left pad)
                FoldedConv v;
                int pad_rows = 0;
                const int cs[1] = {8}, oi[1] = {0};
                rc = fold_same_conv(fold_conv(dsW[0], nullptr, dsBN[0], 8, 8, 7, 8), 8, cs, oi, 1, 3, v, pad_rows);
                if (rc == VP_OK) rc = build(ts.ds0f, v, PN_FOLD * 8, PN_FOLD * 8, 3, 1, pad_rows);
                if (rc != VP_OK) return rc;
            }
            last = f;
            if (i == 4) break;
            // stride-4 conv on the [T / 4][4 f] view: W2[co][q f + ci][tap] = w[co][ci][4 tap + q]
            const FoldedConv fd = fold_conv(ddW[i], nullptr, ddBN[i], f, f, 7, f);
            FoldedConv v;
            v.b = fd.b;
            v.w.assign((size_t)f * 4 * f * 2, 0.f);
            for (int co = 0; co < f; ++co)
                for (int ci = 0; ci < f; ++ci)
                    for (int j = 0; j < 7; ++j)
                        v.w[((size_t)co * 4 * f + (j & 3) * f + ci) * 2 + (j >> 2)] = fd.w[((size_t)co * f + ci) * 7 + j];
            rc = build(ts.dd[i], v, 4 * f, f, 2, f == 64 ? g_dd : 1, 0);
            if (rc != VP_OK) return rc;
        }
        for (int i = 0; i < 4; ++i) {
            const int f = kPnC[3 - i];
            // ConvTranspose1d W (cin = last, cout = f, 7) + BN -> k = 2 conv with 4 f columns: row r, column q f + co
            std::vector<double> sc, sh;
            bn_affine(utBN[i], sc, sh);
            FoldedConv v;
            v.w.assign((size_t)4 * f * last * 2, 0.f);
            v.b.assign((size_t)4 * f, 0.f);
            for (int q = 0; q < 4; ++q)
                for (int co = 0; co < f; ++co) {
                    v.b[(size_t)q * f + co] = (float)sh[co];
                    for (int ci = 0; ci < last; ++ci) {
                        const float *w7 = utW[i] + ((size_t)ci * f + co) * 7;
                        float *dst = &v.w[(((size_t)q * f + co) * last + ci) * 2];
                        dst[1] = (float)((double)w7[q] * sc[co]);                  // x[r]
                        dst[0] = q < 3 ? (float)((double)w7[q + 4] * sc[co]) : 0.f;  // x[r - 1]
                    }
                }
            rc = build(ts.ut[i], v, last, 4 * f, 2, 4 * f > 128 ? g_ut : 1, 1);
            if (rc != VP_OK) return rc;
            last = f;
            rc = build(ts.us[i], fold_conv(usW[i], nullptr, usBN[i], f, 2 * f, 7, 2 * f), 2 * f, f, 7, f == 64 ? g_us : 1, 3);
            if (rc != VP_OK) return rc;
            if (i == 3) {  // folded form: [skip_0 at offset 3 | up at offset 1 + off = 2] -> head, output offset 3 (3 row taps each)
                FoldedConv v;
                int pad_rows = 0;
                const int cs[2] = {8, 8}, oi[2] = {3, 2};
                rc = fold_same_conv(fold_conv(usW[3], nullptr, usBN[3], 8, 16, 7, 16), 8, cs, oi, 2, 3, v, pad_rows);
                if (rc == VP_OK) rc = build(ts.us3f, v, PN_FOLD * 16, PN_FOLD * 8, 3, 1, pad_rows);
                if (rc != VP_OK) return rc;
            }
        }
    }
    return VP_OK;
}

static int build_pn(vp_model *m, Cursor &cur, Packed &pk) {
    const float *incW = cur.take(8 * 3 * 7), *incB = cur.take(8);
    const BN incBN = take_bn(cur, 8);
    m->inc = pack_conv(pk, incW, incB, &incBN, 8, 3, 7);
    const float *dsW[5], *ddW[4], *utW[4], *usW[4];
    BN dsBN[5], ddBN[4], utBN[4], usBN[4];
    int last = 8;
    for (int i = 0; i < 5; ++i) {
        const int f = kPnC[i];
        dsW[i] = cur.take((int64_t)f * last * 7);
        dsBN[i] = take_bn(cur, f);
        m->down_same[i] = pack_conv(pk, dsW[i], nullptr, &dsBN[i], f, last, 7);
        last = f;
        if (i < 4) {
            ddW[i] = cur.take((int64_t)f * f * 7);
            ddBN[i] = take_bn(cur, f);
            m->down_down[i] = pack_conv(pk, ddW[i], nullptr, &ddBN[i], f, f, 7);
        }
    }
    for (int i = 0; i < 4; ++i) {
        const int f = kPnC[3 - i];
        utW[i] = cur.take((int64_t)last * f * 7);
        utBN[i] = take_bn(cur, f);
        m->up_t[i] = pack_convt(pk, utW[i], utBN[i], last, f);
        last = f;
        usW[i] = cur.take((int64_t)f * 2 * f * 7);
        usBN[i] = take_bn(cur, f);
        m->up_same[i] = pack_conv(pk, usW[i], nullptr, &usBN[i], f, 2 * f, 7);
    }
    const float *W = cur.take(3 * 8), *b = cur.take(3);
    m->outc = pack_conv(pk, W, b, nullptr, 3, 8, 1);
    std::memcpy(m->pn_head_w, W, sizeof(m->pn_head_w));
    std::memcpy(m->pn_head_b, b, sizeof(m->pn_head_b));
    {
        const FoldedConv f = fold_conv(incW, incB, incBN, 8, 3, 7, 3);
        std::memcpy(m->pn_inc_w, f.w.data(), sizeof(m->pn_inc_w));
        std::memcpy(m->pn_inc_b, f.b.data(), sizeof(m->pn_inc_b));
    }
    if (cur.left == 0) {
        int rc = build_pn_tc(m, incW, incB, incBN, dsW, dsBN, ddW, ddBN, utW, utBN, usW, usBN);
        if (rc != VP_OK) return rc;
    }
    m->tap_names =
        "inc,down0_same,down0_down,down1_same,down1_down,down2_same,down2_down,down3_same,down3_down,down4_same,"
        "up0_cat,up0_same,up1_cat,up1_same,up2_cat,up2_same,up3_cat,up3_same";
    return VP_OK;
}

// ---- forward plan ---------------------------------------------------------------------------------
struct Arena {
    char *base;
    int64_t used, cap;
    float *take(int64_t n_floats) {
        const int64_t off = align_up(used, 256);
        used = off + n_floats * (int64_t)sizeof(float);
        return base ? reinterpret_cast<float *>(base + off) : nullptr;
    }
    int check() const {
        if (used <= cap) return VP_OK;
        set_error("forward workspace too small: need %lld bytes, have %lld", (long long)used, (long long)cap);
        return VP_ERR_WORKSPACE;
    }
};

struct Tap {
    const char *name;
    const float *ptr;
    int64_t floats;
};

struct Runner {
    vp_model *m;
    cudaStream_t s;
    int B;
    int precision = VP_PREC_FP32;
    int keep_lo = 0, keep_hi = 1 << 30;  // output samples per window that must be computed (the rest is blinded by the caller)
    // when set, the windows are cut from this trace by the fused slicer + encoder.convs.0 kernel (tensor-core path)
    const void *trace = nullptr;
    int trace_dtype = 0, peak_scope = 0, taper = 0;
    int64_t ch_stride = 0;
    const int64_t *starts = nullptr;
    bool dry;               // only measure the workspace
    const char *stop_name;  // stop after this tap (debug)
    Tap hit{nullptr, nullptr, 0};
    bool stopped = false;
    int rc = VP_OK;

    const float *W(int64_t off) const { return m->d_weights + off; }

    bool tap(const char *name, const float *ptr, int64_t floats) {
        if (stop_name && !stopped && std::strcmp(stop_name, name) == 0) {
            hit = Tap{name, ptr, floats};
            stopped = true;
        }
        return stopped;
    }
    bool go() const { return !dry && !stopped && rc == VP_OK; }

    // generic conv step
    void conv(const ConvW &cw, int stride, int ups, int pool, int act, const AffW *pre, const float *res, int64_t r_bs,
              const float *x, int64_t x_bs, int64_t x_gs, int Lin, int Lin_eff, int pad_left, int Lconv, int Lout,
              float *y, int64_t y_bs, int64_t y_gs, int G, int64_t w_gs, int64_t b_gs) {
        if (!go()) return;
        ConvKey key{cw.cin, cw.coutp, cw.k, stride, ups, pool, act, pre ? 1 : 0, res ? 1 : 0};
        conv_launch_fn fn = find_conv_fp32(key, Lconv);
        if (!fn) {
            set_error("no fp32 conv instance for cin=%d coutp=%d k=%d stride=%d ups=%d pool=%d act=%d pre=%d res=%d",
                      key.cin, key.coutp, key.k, key.stride, key.ups, key.pool, key.act, key.pre, key.res);
            rc = VP_ERR_UNSUPPORTED;
            return;
        }
        ConvP p;
        p.x = x;
        p.x_bs = x_bs;
        p.x_gs = x_gs;
        p.w = W(cw.w);
        p.w_gs = w_gs;
        p.bias = W(cw.b);
        p.b_gs = b_gs;
        p.pre_scale = pre ? W(pre->scale) : nullptr;
        p.pre_shift = pre ? W(pre->shift) : nullptr;
        p.res = res;
        p.r_bs = r_bs;
        p.y = y;
        p.y_bs = y_bs;
        p.y_gs = y_gs;
        p.Lin = Lin;
        p.Lin_eff = Lin_eff;
        p.Lconv = Lconv;
        p.Lout = Lout;
        p.pad_left = pad_left;
        p.cout_store = cw.cout;
        rc = fn(p, B, G, s);
    }
};

static int run_eqt(Runner &r, const float *x, float *y, Arena &ar) {
    vp_model *m = r.m;
    const int64_t B = r.B;
    const int L = m->in_samples;
    // lengths through the encoder
    int len[8];
    len[0] = L;
    for (int i = 0; i < 7; ++i) len[i + 1] = (len[i] + (len[i] & 1)) / 2;
    const int T = len[7];
    // decoder lengths (SeisBench Decoder crops: drop the last sample where the encoder padded)
    int dlen[8], crop[7];
    dlen[0] = T;
    for (int i = 0; i < 7; ++i) {
        crop[i] = (len[6 - i] & 1) ? 1 : 0;
        dlen[i + 1] = 2 * dlen[i] - crop[i];
    }
    if (dlen[7] != L) {
        set_error("EQTransformer: decoder length %d != in_samples %d", dlen[7], L);
        return VP_ERR_ARG;
    }
    // workspace
    int64_t big = 0;
    for (int i = 0; i < 7; ++i) {
        big = std::max<int64_t>(big, (int64_t)kEncC[i + 1] * len[i + 1]);
        big = std::max<int64_t>(big, 3 * (int64_t)kDecC[i + 1] * dlen[i + 1]);
    }
    const bool tc = r.precision != VP_PREC_FP32;
    const int split = (r.precision == VP_PREC_F16X3) ? 2 : 1;
    vp_model::TcSet &ts = m->tc[split == 2 ? 0 : 1];
    // 16-bit channel-last ping-pong buffers of the tensor-core path: [split][B * big16] elements each
    const int64_t big16 = std::max<int64_t>(big, (int64_t)8 * L);
    float *P = nullptr, *Q = nullptr;
    uint16_t *P16 = nullptr, *Q16 = nullptr;
    float *hbuf = nullptr;  // tensor-core path: last decoder stage in fp32 (3, B, 8, L) for the CUDA-core heads
    if (tc) {
        P16 = reinterpret_cast<uint16_t *>(ar.take((B * big16 * split + 1) / 2));
        Q16 = reinterpret_cast<uint16_t *>(ar.take((B * big16 * split + 1) / 2));
        hbuf = ar.take(3 * B * 8 * (int64_t)L);
    } else {
        P = ar.take(B * big);
        Q = ar.take(B * big);
    }
    const int64_t split16 = B * big16;  // elements between the hi and lo planes
    float *r0 = ar.take(B * 64 * T), *r1 = ar.take(B * 64 * T), *r2 = ar.take(B * 64 * T);
    float *xres = tc ? ar.take(B * 64 * T) : nullptr;  // tensor-core path: residual stream, fp32 row-major [B][T][64]
    float *lproj = tc ? ar.take(B * 128 * T) : nullptr;  // tensor-core path: BiLSTM-0 input projection, fp32 [B][T][2][16][4]
    static const bool lproj_off = getenv("VP_LSTM_PROJ") && atoi(getenv("VP_LSTM_PROJ")) == 0;  // debugging aid
    const bool use_lproj = tc && !lproj_off && r.stop_name == nullptr;  // the debug taps keep the fp32 (B, 64, T) stack output
    float *lo = ar.take(B * 32 * T);
    float *s0 = ar.take(B * 16 * T), *s1 = ar.take(B * 16 * T);
    float *din = ar.take(3 * B * 16 * T);  // [group][B][16][T]: decoder inputs
    float *plo = ar.take(2 * B * 16 * T);  // pick LSTM outputs [group][B][16][T]
    if (r.dry) return VP_OK;
    if (int rc = ar.check()) return rc;  // every buffer is laid out: refuse an undersized workspace BEFORE any launch
    if (tc && !ts.ready) {
        set_error("tensor-core weight set is not available");
        return VP_ERR_UNSUPPORTED;
    }

    float *pp[2] = {P, Q};
    uint16_t *pp16[2] = {P16, Q16};
    static const char *enc_names[7] = {"enc0", "enc1", "enc2", "enc3", "enc4", "enc5", "enc6"};
    if (!tc) {
        // ---- encoder (fp32 CUDA cores)
        const float *cur = x;
        int cur_c = 3;
        for (int i = 0; i < 7; ++i) {
            float *dst = (i == 6) ? r0 : pp[i & 1];
            const ConvW &cw = m->enc[i];
            r.conv(cw, 1, 1, 2, ACT_RELU, nullptr, nullptr, 0, cur, (int64_t)cur_c * len[i], 0, len[i], len[i], cw.k / 2,
                   len[i], len[i + 1], dst, (int64_t)cw.cout * len[i + 1], 0, 1, 0, 0);
            if (r.tap(enc_names[i], dst, B * cw.cout * len[i + 1])) return r.rc;
            cur = dst;
            cur_c = cw.cout;
        }
    } else {
        // ---- encoder on the tensor cores: x (B,3,L) fp32 -> channel-last 16-bit [B][L][8] -> 7 x (conv, ReLU, pool)
        const uint16_t *cur16 = Q16;
        int first_layer = 0;
        if (r.trace) {  // K1 + encoder.convs.0 + ReLU + pool on the CUDA cores, straight from the record
            if (r.go())
                r.rc = launch_slice_enc0(r.trace, r.trace_dtype, r.ch_stride, r.starts, B, L, r.peak_scope, r.taper, m->enc0_w,
                                         m->enc0_b, split, P16, split16, r.s);
            cur16 = P16;
            first_layer = 1;
        } else if (r.go()) {
            r.rc = launch_pack_cl16(x, 3 * (int64_t)L, L, (int)B, 3, L, split, Q16, split16, 1, r.s);
        }
        const bool enc6_tap = r.stop_name && std::strcmp(r.stop_name, "enc6") == 0;
        for (int i = first_layer; i < 7; ++i) {
            const TcLayer &tl = ts.enc[i];
            if (i == 1 && ts.enca.ready && len[1] == 3000) {  // convs.1-3 in one kernel (fused_enc.cu); VP_ENC_FUSED=0: layer by layer
                const char *ef = getenv("VP_ENC_FUSED");
                if (!(ef && atoi(ef) == 0)) {
                    if (r.go()) r.rc = enca_launch(ts.enca, cur16, split16, (int)B, pp16[1], split16, r.s);
                    cur16 = pp16[1];
                    i = 3;
                    continue;
                }
            }
            if (r.go()) {
                TcIO io;
                io.x = cur16;
                io.x_split = split16;
                io.x_gs = 0;
                io.T_in = len[i];
                io.NS = (int)B;
                io.w_dev = ts.d_w + tl.w_off;
                io.b_dev = ts.d_b + tl.b_off;
                io.act = ACT_RELU;
                io.pool = 2;
                static const bool encfold_off = getenv("VP_ENC_FOLD") && atoi(getenv("VP_ENC_FOLD")) == 0;  // debugging aid
                if ((i == 1 || i == 2) && !encfold_off && len[i] % 4 == 0) {
                    // encoder.convs.1 / .2 on the [T / 4][4 C] view: conv + ReLU + MaxPool inside the accumulator row
                    const TcLayer &tf = ts.encf[i];
                    io.T_in = len[i] / 4;
                    io.w_dev = ts.d_w + tf.w_off;
                    io.b_dev = ts.d_b + tf.b_off;
                    io.pool = 1;
                    io.foldpool = 1;
                    io.out_fmt = 0;
                    io.y = pp16[i & 1];
                    io.y_split = split16;
                    io.y_gs = 0;
                    io.y_ss = 0;
                    io.y_cs = 0;
                    io.cout_cl = tf.cout;
                    r.rc = tc_launch(tf, io, r.s);
                    cur16 = pp16[i & 1];
                    continue;
                }
                if (i == 6 && !enc6_tap) {
                    // last encoder stage feeds the res-CNN stack: fp32 residual stream (row-major) + the 16-bit
                    // relu(bn1(x)) operand of res_cnn_stack.members.0.conv1
                    io.out_fmt = 0;
                    io.y = pp16[0];
                    io.y_split = split16;
                    io.y_gs = 0;
                    io.y_ss = 0;
                    io.y_cs = 0;
                    io.cout_cl = 64;
                    io.y32 = xres;
                    io.post_scale = r.W(m->res[0].n1.scale);
                    io.post_shift = r.W(m->res[0].n1.shift);
                } else if (i == 6) {  // debug tap: (B, 64, T) channel-first fp32
                    io.out_fmt = 1;
                    io.y = r0;
                    io.y_split = 0;
                    io.y_gs = 0;
                    io.y_ss = 64 * (int64_t)T;
                    io.y_cs = T;
                    io.cout_cl = 0;
                } else {
                    io.out_fmt = 0;
                    io.y = pp16[i & 1];
                    io.y_split = split16;
                    io.y_gs = 0;
                    io.y_ss = 0;
                    io.y_cs = 0;
                    io.cout_cl = tl.cout;
                }
                r.rc = tc_launch(tl, io, r.s);
            }
            if (i == 6 && r.tap(enc_names[i], r0, B * 64 * T)) return r.rc;
            cur16 = pp16[i & 1];
        }
    }
    // ---- res-CNN stack: x in ra; tmp r1; out rb
    static const char *res_names[7] = {"res0", "res1", "res2", "res3", "res4", "res5", "res6"};
    float *ra = r0, *rb = r2;
    if (tc) {
        // tensor cores: conv1 reads relu(bn1(x)) (16-bit, written by the previous epilogue) and writes relu(bn2(y));
        // conv2 adds the fp32 residual stream in place and writes the next block's relu(bn1(x))
        static const bool resfuse_off = getenv("VP_FUSED_RES") && atoi(getenv("VP_FUSED_RES")) == 0;  // debugging aid
        const bool fused_res = use_lproj && !resfuse_off && T + 1 <= 64;
        static const bool res_v1 = getenv("VP_RES_V1") && atoi(getenv("VP_RES_V1")) != 0;  // A/B aid: the layer-by-layer-in-L2 variant
        if (fused_res && !res_v1 && r.go()) {  // all 14 convs with the activations on chip (fused_res2.cu)
            ResStack2P sp;
            std::memset(&sp, 0, sizeof(sp));
            sp.x = pp16[0];
            sp.xres = xres;
            sp.y = pp16[0];  // tiles are read completely (TMA) before their stack output is stored, and tiles do not overlap
            sp.split16 = split16;
            for (int i = 0; i < 7; ++i) {
                sp.w[2 * i] = ts.d_w + ts.res1[i].w_off;
                sp.w[2 * i + 1] = ts.d_w + ts.res2[i].w_off;
                sp.ntaps[2 * i] = sp.ntaps[2 * i + 1] = kResK[i];
            }
            sp.par = ts.d_res_par;
            sp.NS = (int)B;
            sp.T = T;
            sp.fmt16 = split == 2 ? 0 : 1;
            r.rc = resstack2_launch(sp, split, r.s);
        } else if (fused_res && r.go()) {  // all 14 convs in one persistent launch, activations through L2 (fused_res.cu)
            ResStackP sp;
            std::memset(&sp, 0, sizeof(sp));
            sp.n_layers = 14;
            sp.NS = (int)B;
            sp.T = T;
            sp.fmt16 = split == 2 ? 0 : 1;
            sp.split16 = split16;
            for (int i = 0; i < 7; ++i) {
                ResLayerP &a = sp.l[2 * i], &b = sp.l[2 * i + 1];
                a.x = pp16[0];
                a.y = pp16[1];
                a.w = ts.d_w + ts.res1[i].w_off;
                a.bias = ts.d_b + ts.res1[i].b_off;
                a.psc = r.W(m->res[i].n2.scale);
                a.psh = r.W(m->res[i].n2.shift);
                a.res = nullptr;
                a.ntaps = kResK[i];
                a.affine = 1;
                a.write_res = 0;
                b.x = pp16[1];
                b.y = pp16[0];
                b.w = ts.d_w + ts.res2[i].w_off;
                b.bias = ts.d_b + ts.res2[i].b_off;
                b.res = xres;
                b.ntaps = kResK[i];
                if (i < 6) {
                    b.psc = r.W(m->res[i + 1].n1.scale);
                    b.psh = r.W(m->res[i + 1].n1.shift);
                    b.affine = 1;
                    b.write_res = 1;
                } else {  // stack output as the 16-bit operand of the BiLSTM input-projection GEMM
                    b.psc = b.psh = nullptr;
                    b.affine = 0;
                    b.write_res = 0;
                }
            }
            // optional sub-chunks (the working set, 36 KB per window, then fits the 126 MB L2): no gain measured, the layer sync costs as much
            static const int res_sub = getenv("VP_RES_SUB") ? atoi(getenv("VP_RES_SUB")) : 0;  // measured: 0 (whole chunk) 8.32 ms, 2048 8.35 ms, 1024 8.89 ms
            const int64_t sub = res_sub > 0 ? (res_sub + 1) / 2 * 2 : B;  // even: tiles hold two sequences
            for (int64_t b0 = 0; b0 < B && r.rc == VP_OK; b0 += sub) {
                ResStackP sq = sp;
                sq.NS = (int)std::min<int64_t>(sub, B - b0);
                for (int l = 0; l < sq.n_layers; ++l) {
                    sq.l[l].x += b0 * T * 64;
                    sq.l[l].y += b0 * T * 64;
                    if (sq.l[l].res) sq.l[l].res += b0 * T * 64;
                }
                r.rc = resstack_launch(sq, split, r.s);
            }
        }
        for (int i = 0; i < 7 && r.go() && !fused_res; ++i) {
            TcIO io;
            io.x = pp16[0];
            io.x_split = split16;
            io.x_gs = 0;
            io.T_in = T;
            io.NS = (int)B;
            io.w_dev = ts.d_w + ts.res1[i].w_off;
            io.b_dev = ts.d_b + ts.res1[i].b_off;
            io.act = ACT_NONE;
            io.pool = 1;
            io.out_fmt = 0;
            io.y = pp16[1];
            io.y_split = split16;
            io.y_gs = io.y_ss = io.y_cs = 0;
            io.cout_cl = 64;
            io.post_scale = r.W(m->res[i].n2.scale);
            io.post_shift = r.W(m->res[i].n2.shift);
            r.rc = tc_launch(ts.res1[i], io, r.s);
            if (r.rc != VP_OK) break;
            TcIO io2;
            io2.x = pp16[1];
            io2.x_split = split16;
            io2.x_gs = 0;
            io2.T_in = T;
            io2.NS = (int)B;
            io2.w_dev = ts.d_w + ts.res2[i].w_off;
            io2.b_dev = ts.d_b + ts.res2[i].b_off;
            io2.act = ACT_NONE;
            io2.pool = 1;
            io2.res = xres;
            io2.y_gs = 0;
            if (i < 6) {
                io2.out_fmt = 0;
                io2.y = pp16[0];
                io2.y_split = split16;
                io2.y_ss = io2.y_cs = 0;
                io2.cout_cl = 64;
                io2.y32 = xres;
                io2.post_scale = r.W(m->res[i + 1].n1.scale);
                io2.post_shift = r.W(m->res[i + 1].n1.shift);
            } else if (use_lproj) {  // stack output as the 16-bit operand of the BiLSTM input-projection GEMM
                io2.out_fmt = 0;
                io2.y = pp16[0];
                io2.y_split = split16;
                io2.y_ss = io2.y_cs = 0;
                io2.cout_cl = 64;
            } else {  // stack output -> BiLSTM input, fp32 (B, 64, T) channel-first
                io2.out_fmt = 1;
                io2.y = ra;
                io2.y_split = 0;
                io2.y_ss = 64 * (int64_t)T;
                io2.y_cs = T;
                io2.cout_cl = 0;
            }
            r.rc = tc_launch(ts.res2[i], io2, r.s);
        }
        if (use_lproj && r.go()) {  // W_ih x + b of both directions for all time steps: (B T) x 64 x 128 GEMM
            TcIO io;
            io.x = pp16[0];
            io.x_split = split16;
            io.x_gs = 0;
            io.T_in = T;
            io.NS = (int)B;
            io.w_dev = ts.d_w + ts.lproj.w_off;
            io.b_dev = ts.d_b + ts.lproj.b_off;
            io.act = ACT_NONE;
            io.pool = 1;
            io.out_fmt = 3;  // only the fp32 row-major second output
            io.y = nullptr;
            io.y_split = 0;
            io.y_gs = io.y_ss = io.y_cs = 0;
            io.cout_cl = 128;
            io.y32 = lproj;
            r.rc = tc_launch(ts.lproj, io, r.s);
        }
        if (r.tap("res6", ra, B * 64 * T)) return r.rc;
    }
    for (int i = 0; i < 7 && !tc; ++i) {
        const int k = kResK[i];
        const int pad = (k == 3) ? 1 : 0;  // k == 2: right-pad one zero == out-of-range reads as 0
        r.conv(m->res[i].c1, 1, 1, 1, ACT_NONE, &m->res[i].n1, nullptr, 0, ra, 64 * T, 0, T, T, pad, T, T, r1, 64 * T, 0,
               1, 0, 0);
        r.conv(m->res[i].c2, 1, 1, 1, ACT_NONE, &m->res[i].n2, ra, 64 * T, r1, 64 * T, 0, T, T, pad, T, T, rb, 64 * T,
               0, 1, 0, 0);
        if (r.tap(res_names[i], rb, B * 64 * T)) return r.rc;
        std::swap(ra, rb);
    }
    // ---- BiLSTM blocks
    static const char *bl_names[3][2] = {{"bilstm0_lstm", "bilstm0"}, {"bilstm1_lstm", "bilstm1"}, {"bilstm2_lstm", "bilstm2"}};
    const float *seq = ra;
    int seq_c = 64;
    float *sp[2] = {s0, s1};
    for (int i = 0; i < 3; ++i) {
        if (r.go()) {
            LstmP p;
            p.x = seq;
            p.x_bs = (int64_t)seq_c * T;
            p.x_gs = 0;
            p.w_ih = r.W(m->bil[i].lstm.ih);
            p.w_hh = r.W(m->bil[i].lstm.hh);
            p.bias = r.W(m->bil[i].lstm.b);
            p.w_gs_ih = p.w_gs_hh = p.w_gs_b = 0;
            p.y = lo;
            p.y_bs = 32 * T;
            p.y_gs = 0;
            p.T = T;
            p.ndir = 2;
            p.B = (int)B;
            if (i == 0 && use_lproj) p.proj = lproj;
            r.rc = launch_lstm(p.proj ? 0 : seq_c, p, 1, r.s);
        }
        if (r.tap(bl_names[i][0], lo, B * 32 * T)) return r.rc;
        float *dst = sp[i & 1];
        r.conv(m->bil[i].conv, 1, 1, 1, ACT_NONE, nullptr, nullptr, 0, lo, 32 * T, 0, T, T, 0, T, T, dst, 16 * T, 0, 1, 0,
               0);
        if (r.tap(bl_names[i][1], dst, B * 16 * T)) return r.rc;
        seq = dst;
        seq_c = 16;
    }
    // ---- transformers: seq (s0 after 3 blocks) -> s1 -> din[0]
    {
        const float *tin = seq;
        float *tout[2] = {(seq == s0) ? s1 : s0, din};
        static const char *tn[2] = {"transformer_d0", "transformer_d"};
        for (int i = 0; i < 2; ++i) {
            if (r.go()) {
                AttnP p;
                p.x = tin;
                p.x_bs = 16 * T;
                p.x_gs = 0;
                p.w = r.W(m->tr[i]);
                p.w_gs = 0;
                p.y = tout[i];
                p.y_bs = 16 * T;
                p.y_gs = 0;
                p.T = T;
                p.B = (int)B;
                p.width = 0;
                p.mode = 0;
                r.rc = launch_attention(p, 1, r.s);
            }
            if (r.tap(tn[i], tout[i], B * 16 * T)) return r.rc;
            tin = tout[i];
        }
    }
    // ---- pick branches: uni-LSTM -> banded attention, both branches as 2 groups
    if (r.go()) {
        LstmP p;
        p.x = din;
        p.x_bs = 16 * T;
        p.x_gs = 0;  // both read the shared encoder output
        p.w_ih = r.W(m->pick_lstm[0].ih);
        p.w_hh = r.W(m->pick_lstm[0].hh);
        p.bias = r.W(m->pick_lstm[0].b);
        p.w_gs_ih = m->pick_lstm[1].ih - m->pick_lstm[0].ih;
        p.w_gs_hh = m->pick_lstm[1].hh - m->pick_lstm[0].hh;
        p.w_gs_b = m->pick_lstm[1].b - m->pick_lstm[0].b;
        p.y = plo;
        p.y_bs = 16 * T;
        p.y_gs = B * 16 * T;
        p.T = T;
        p.ndir = 1;
        p.B = (int)B;
        r.rc = launch_lstm(16, p, 2, r.s);
    }
    if (r.tap("pick_lstm", plo, 2 * B * 16 * T)) return r.rc;
    if (r.go()) {
        AttnP p;
        p.x = plo;
        p.x_bs = 16 * T;
        p.x_gs = B * 16 * T;
        p.w = r.W(m->pick_attn[0]);
        p.w_gs = m->pick_attn[1] - m->pick_attn[0];
        p.y = din + B * 16 * T;
        p.y_bs = 16 * T;
        p.y_gs = B * 16 * T;
        p.T = T;
        p.B = (int)B;
        p.width = 3;
        p.mode = 1;
        r.rc = launch_attention(p, 2, r.s);
    }
    if (r.tap("pick_attn", din + B * 16 * T, 2 * B * 16 * T)) return r.rc;
    if (tc) {
        // ---- decoders + heads on the tensor cores: din (3, B, 16, T) fp32 -> [3][B][T][16] 16-bit
        if (r.go()) r.rc = launch_pack_cl16(din, 16 * (int64_t)T, T, (int)(3 * B), 16, T, split, Q16, split16, 2, r.s);
        const uint16_t *cur16 = Q16;
        static const bool fused_off = getenv("VP_FUSED") && atoi(getenv("VP_FUSED")) == 0;
        const bool fused = ts.decb.ready && !fused_off;
        static const bool deca_off = getenv("VP_FUSED_A") && atoi(getenv("VP_FUSED_A")) == 0;
        const bool fused_a = fused && ts.deca.ready && !deca_off;
        for (int i = 0; i < 7; ++i) {
            const TcLayer &tl = ts.dec[i];
            uint16_t *dst = pp16[i & 1];
            if (fused_a && i == 1) {  // decoder.convs.1 + convs.2 in one kernel: (3, B, 94, 64) -> (3, B, 375, 32)
                if (r.go())
                    r.rc = deca_launch(ts.deca, cur16, split16, B * (int64_t)64 * dlen[1], (int)B, pp16[1], split16, B * (int64_t)32 * dlen[3], r.s);
                cur16 = pp16[1];
                i = 2;  // continue with decoder.convs.3 (the fused tail)
                continue;
            }
            if (fused && i == 3) {  // decoder.convs.3-6 + heads in one kernel, activations in shared memory
                // A/B arm (read per launch so that a test can compare the two in one process): the shared-memory-operand kernel
                const char *v1 = getenv("VP_DECB_V1");
                const bool decb_v1 = v1 && atoi(v1) != 0;
                if (r.go())
                    r.rc = (ts.decb2.ready && !decb_v1)
                               ? decb2_launch(ts.decb2, cur16, split16, B * (int64_t)tl.cin * dlen[i], (int)B, y, r.keep_lo, r.keep_hi, r.s)
                               : decb_launch(ts.decb, cur16, split16, B * (int64_t)tl.cin * dlen[i], (int)B, y, r.keep_lo, r.keep_hi, r.s);
                return r.rc;
            }
            if (r.go()) {
                TcIO io;
                io.x = cur16;
                io.x_split = split16;
                io.x_gs = B * (int64_t)tl.cin * dlen[i];
                io.T_in = dlen[i];
                io.NS = (int)B;
                io.w_dev = ts.d_w + tl.w_off;
                io.b_dev = ts.d_b + tl.b_off;
                io.act = ACT_RELU;
                io.pool = 1;
                if (i == 6) {  // N = 1 heads run on the CUDA cores: hand them fp32 channel-first activations
                    io.out_fmt = 1;
                    io.y = hbuf;
                    io.y_split = 0;
                    io.y_gs = B * 8 * (int64_t)L;
                    io.y_ss = 8 * (int64_t)L;
                    io.y_cs = L;
                    io.cout_cl = 0;
                } else {
                    io.out_fmt = 0;
                    io.y = dst;
                    io.y_split = split16;
                    io.y_gs = B * (int64_t)tl.cout * dlen[i + 1];
                    io.y_ss = 0;
                    io.y_cs = 0;
                    io.cout_cl = tl.cout;
                }
                r.rc = tc_launch(tl, io, r.s);
            }
            cur16 = dst;
        }
        {
            const ConvW &cw = m->head[0];
            const int64_t w_gs = m->head[1].w - m->head[0].w, b_gs = m->head[1].b - m->head[0].b;
            r.conv(cw, 1, 1, 1, ACT_SIGMOID, nullptr, nullptr, 0, hbuf, 8 * (int64_t)L, B * 8 * (int64_t)L, L, L, 5, L, L, y,
                   3 * (int64_t)L, L, 3, w_gs, b_gs);
        }
        return r.rc;
    }
    // ---- the three decoders as 3 groups per layer
    static const char *dec_names[7] = {"dec0", "dec1", "dec2", "dec3", "dec4", "dec5", "dec6"};
    const float *dcur = din;
    for (int i = 0; i < 7; ++i) {
        float *dst = pp[i & 1];
        const ConvW &cw = m->dec[0][i];
        const int64_t w_gs = m->dec[1][i].w - m->dec[0][i].w, b_gs = m->dec[1][i].b - m->dec[0][i].b;
        r.conv(cw, 1, 2, 1, ACT_RELU, nullptr, nullptr, 0, dcur, (int64_t)cw.cin * dlen[i], B * cw.cin * dlen[i], dlen[i],
               dlen[i + 1], cw.k / 2, dlen[i + 1], dlen[i + 1], dst, (int64_t)cw.cout * dlen[i + 1],
               B * cw.cout * dlen[i + 1], 3, w_gs, b_gs);
        if (r.tap(dec_names[i], dst, 3 * B * cw.cout * dlen[i + 1])) return r.rc;
        dcur = dst;
    }
    // ---- heads: sigmoid(conv k11) -> y (B, 3, L)
    {
        const ConvW &cw = m->head[0];
        const int64_t w_gs = m->head[1].w - m->head[0].w, b_gs = m->head[1].b - m->head[0].b;
        r.conv(cw, 1, 1, 1, ACT_SIGMOID, nullptr, nullptr, 0, dcur, 8 * (int64_t)L, B * 8 * (int64_t)L, L, L, 5, L, L, y,
               3 * (int64_t)L, L, 3, w_gs, b_gs);
    }
    return r.rc;
}

// PhaseNet on the tensor cores.  Buffers are 16-bit channel-last [split][B][rows][C]:
//   Bk  dense activations,  Al  skip_l stored at row offset padl[l] inside a zero-padded pitch of 4 * (len[l+1] + 1) rows
//   (the stride-4 conv reads it as [rows / 4][4 C]),  Ul  ConvTranspose1d output, written as [Lin + 1][4 C] = [4 (Lin + 1)][C].
// The channel concatenation [skip | up] is never materialised: the last conv of an up level stages its tile from both
// buffers (TcIO::x2), the crop [1:-2] and the centring offset being row offsets of the second source.
static int run_pn_tc(Runner &r, const float *x, float *y, Arena &ar) {
    vp_model *m = r.m;
    const int64_t B = r.B;
    const int L0 = m->in_samples;
    const int split = (r.precision == VP_PREC_F16X3) ? 2 : 1;
    vp_model::PnTcSet &ts = m->pn_tc[split == 2 ? 0 : 1];
    static const int padl[4] = {3, 2, 1, 2}, padr[4] = {3, 3, 3, 3};
    int len[5], pitchA[4];
    len[0] = L0;
    for (int i = 0; i < 4; ++i) {
        len[i + 1] = (len[i] + padl[i] + padr[i] - 7) / 4 + 1;
        pitchA[i] = 4 * (len[i + 1] + 1);
        if (pitchA[i] < padl[i] + len[i]) {
            set_error("PhaseNet: padded pitch %d too small for level %d", pitchA[i], i);
            return VP_ERR_ARG;
        }
    }
    auto take16 = [&](int64_t rows, int c) { return reinterpret_cast<uint16_t *>(ar.take((B * rows * c * split + 1) / 2)); };
    static const bool fold_off = getenv("VP_PN_FOLD") && atoi(getenv("VP_PN_FOLD")) == 0;  // debugging aid
    const bool fold = !fold_off;
    const int pitchB0 = fold ? (L0 + PN_FOLD - 1) / PN_FOLD * PN_FOLD : L0;  // folded reads need whole rows of 4 samples
    uint16_t *X16 = take16(L0, 8), *B0 = take16(pitchB0, 8);
    uint16_t *A[4], *Bd[5], *U[4], *Bu[4];
    for (int i = 0; i < 4; ++i) {
        A[i] = take16(pitchA[i], kPnC[i]);
        Bd[i] = take16(len[i + 1], kPnC[i]);
    }
    Bd[4] = take16(len[4], 128);
    for (int i = 0; i < 4; ++i) {
        const int lvl = 3 - i;
        U[i] = take16(pitchA[lvl], kPnC[lvl]);  // 4 * (len[lvl + 1] + 1) rows
        Bu[i] = take16(len[lvl], kPnC[lvl]);
    }
    float *tapbuf = ar.take(B * 8 * (int64_t)L0);  // debug taps: any layer as fp32 (B, C, T), C * T <= 8 * L0
    if (r.dry) return VP_OK;
    if (int rc = ar.check()) return rc;  // every buffer is laid out: refuse an undersized workspace BEFORE any launch
    if (!ts.ready) {
        set_error("tensor-core weight set is not available");
        return VP_ERR_UNSUPPORTED;
    }
    auto base_io = [&](const TcLayer &tl, const uint16_t *src, int rows_in, int c_in, int T_in) {
        TcIO io;
        io.x = src;
        io.x_split = B * (int64_t)rows_in * c_in;
        io.x_gs = 0;
        io.T_in = T_in;
        io.NS = (int)B;
        io.w_dev = ts.d_w + tl.w_off;
        io.b_dev = ts.d_b + tl.b_off;
        io.act = ACT_RELU;
        io.pool = 1;
        io.out_fmt = 0;
        io.y = nullptr;
        io.y_split = 0;
        io.y_gs = 0;
        io.y_ss = 0;
        io.y_cs = 0;
        io.cout_cl = 0;
        io.x_pitch = rows_in;
        io.ca = 1;  // PhaseNet: L1-cached staging loads (see TcIO::ca)
        return io;
    };
    // dense 16-bit output [B][rows_out][c_total] (+ group column offset), or the fp32 debug tap
    auto run_layer = [&](const char *name, const TcLayer &tl, TcIO io, uint16_t *dst, int rows_out, int y_roff, int c_total,
                         int T_store) -> bool {
        const bool tapped = r.stop_name && !r.stopped && name && std::strcmp(r.stop_name, name) == 0;
        if (tapped) {
            const int c_tap = io.fold > 1 ? io.fold_c : c_total;  // folded layers: channels of one sample
            io.out_fmt = 1;
            io.y = tapbuf;
            io.y_split = 0;
            io.y_gs = (int64_t)tl.cout * T_store;  // group g holds channels [g * cout, (g + 1) * cout)
            io.y_ss = (int64_t)c_tap * T_store;
            io.y_cs = T_store;
            io.cout_cl = 0;
            c_total = c_tap;
        } else {
            io.y = dst;
            io.y_split = B * (int64_t)rows_out * c_total;
            io.y_gs = tl.cout;
            io.cout_cl = c_total;
            io.y_pitch = rows_out;
            io.y_roff = y_roff;
        }
        if (r.go()) r.rc = tc_launch(tl, io, r.s);
        if (tapped) r.tap(name, tapbuf, B * c_total * T_store);
        return r.stopped || r.rc != VP_OK;
    };
    if (r.trace) {  // K1 + inc + in_bn + ReLU on the CUDA cores, straight from the record (the fp32 windows never exist)
        if (r.go())
            r.rc = launch_slice_enc0(r.trace, r.trace_dtype, r.ch_stride, r.starts, B, L0, r.peak_scope, r.taper, m->pn_inc_w,
                                     m->pn_inc_b, split, B0, B * (int64_t)pitchB0 * 8, r.s, 7, pitchB0);
    } else {
        if (r.go()) r.rc = launch_pack_cl16(x, 3 * (int64_t)L0, L0, (int)B, 3, L0, split, X16, B * (int64_t)L0 * 8, 1, r.s);
        if (r.go() && pitchB0 > L0) {  // rows past the window inside B0's pitch are conv zero padding for the folded reader
            const cudaError_t e = cudaMemset2DAsync(B0 + (size_t)L0 * 8, (size_t)pitchB0 * 16, 0, (size_t)(pitchB0 - L0) * 16, (size_t)B * split, r.s);
            if (e != cudaSuccess) {
                set_error("PhaseNet: clearing the B0 tail failed: %s", cudaGetErrorString(e));
                return VP_ERR_CUDA;
            }
        }
        if (run_layer("inc", ts.inc, base_io(ts.inc, X16, L0, 8, L0), B0, pitchB0, 0, 8, L0)) return r.rc;
    }
    static const char *ds_names[5] = {"down0_same", "down1_same", "down2_same", "down3_same", "down4_same"};
    static const char *dd_names[4] = {"down0_down", "down1_down", "down2_down", "down3_down"};
    static const char *us_names[4] = {"up0_same", "up1_same", "up2_same", "up3_same"};
    const uint16_t *cur = B0;
    int cur_c = 8;
    for (int i = 0; i < 4; ++i) {
        const int f = kPnC[i];
        if (i == 0 && fold) {
            // level 0, folded: every row of A0 is written (samples outside the sequence as zeros), no clearing needed
            TcIO io = base_io(ts.ds0f, cur, pitchB0 / PN_FOLD, PN_FOLD * 8, pitchB0 / PN_FOLD);
            io.T_valid = pitchA[0] / PN_FOLD;
            io.fold = PN_FOLD;
            io.fold_c = 8;
            io.fold_o = padl[0];
            io.fold_T = len[0];
            if (run_layer(ds_names[0], ts.ds0f, io, A[0], pitchA[0] / PN_FOLD, 0, PN_FOLD * 8, len[0])) return r.rc;
        } else if (r.go()) {  // zero rows around skip_i inside its padded pitch (read by the stride-4 conv as conv padding)
            const size_t pitch_b = (size_t)pitchA[i] * f * 2, head_b = (size_t)padl[i] * f * 2;
            const size_t tail_b = (size_t)(pitchA[i] - padl[i] - len[i]) * f * 2;
            cudaError_t e = cudaMemset2DAsync(A[i], pitch_b, 0, head_b, (size_t)B * split, r.s);
            if (e == cudaSuccess && tail_b)
                e = cudaMemset2DAsync(A[i] + (size_t)(padl[i] + len[i]) * f, pitch_b, 0, tail_b, (size_t)B * split, r.s);
            if (e != cudaSuccess) {
                set_error("PhaseNet: clearing the skip padding failed: %s", cudaGetErrorString(e));
                return VP_ERR_CUDA;
            }
        }
        if (!(i == 0 && fold)) {
            TcIO io = base_io(ts.ds[i], cur, i == 0 ? pitchB0 : len[i], cur_c, len[i]);
            if (run_layer(ds_names[i], ts.ds[i], io, A[i], pitchA[i], padl[i], f, len[i])) return r.rc;
        }
        {   // stride-4 conv: A_i seen as [pitch / 4][4 f], k = 2, 'valid': len[i + 1] output rows
            TcIO io = base_io(ts.dd[i], A[i], pitchA[i] / 4, 4 * f, pitchA[i] / 4);
            io.T_valid = len[i + 1];
            if (run_layer(dd_names[i], ts.dd[i], io, Bd[i], len[i + 1], 0, f, len[i + 1])) return r.rc;
        }
        cur = Bd[i];
        cur_c = f;
    }
    if (run_layer(ds_names[4], ts.ds[4], base_io(ts.ds[4], cur, len[4], cur_c, len[4]), Bd[4], len[4], 0, 128, len[4])) return r.rc;
    cur = Bd[4];
    cur_c = 128;
    int cur_len = len[4];
    for (int i = 0; i < 4; ++i) {
        const int lvl = 3 - i;
        const int f = kPnC[lvl];
        const int Ls = len[lvl];
        const int Lt = (cur_len - 1) * 4 + 7 - 3;  // ConvTranspose1d output, cropped [1:-2]
        const int off = (Lt - Ls) / 2;
        if (off < 0 || off + Ls > Lt || 4 * (cur_len + 1) != pitchA[lvl]) {
            set_error("PhaseNet: skip merge mismatch (Lt=%d, Ls=%d)", Lt, Ls);
            return VP_ERR_ARG;
        }
        {   // ConvTranspose1d as a k = 2 conv with 4 f columns: rows 0..cur_len of the [cur_len + 1][4 f] view of U
            TcIO io = base_io(ts.ut[i], cur, cur_len, cur_c, cur_len);
            io.T_valid = cur_len + 1;
            if (i == 3 && fold) {  // the folded reader sees whole rows: samples outside the crop [1 + off, 1 + off + Ls) become zeros
                io.fold = 4;
                io.fold_c = f;
                io.fold_o = 1 + off;
                io.fold_T = Ls;
            }
            if (run_layer(nullptr, ts.ut[i], io, U[i], cur_len + 1, 0, 4 * f, cur_len + 1)) return r.rc;
        }
        {   // conv over [skip | up] without the concatenation
            const bool fo = i == 3 && fold;
            const TcLayer &tu = fo ? ts.us3f : ts.us[i];
            TcIO io = base_io(tu, A[lvl], fo ? pitchA[lvl] / PN_FOLD : pitchA[lvl], fo ? PN_FOLD * f : f, fo ? pitchA[lvl] / PN_FOLD : Ls);
            io.x2 = U[i];
            io.x2_split = B * (int64_t)pitchA[lvl] * f;
            if (fo) {  // offsets (skip 3, up 1 + off = 2, output 3) live in the folded weights
                if (padl[lvl] != 3 || 1 + off != 2) {
                    set_error("PhaseNet: folded level-0 layer built for offsets (3, 2), got (%d, %d)", padl[lvl], 1 + off);
                    return VP_ERR_ARG;
                }
                io.x2_pitch = pitchA[lvl] / PN_FOLD;
                io.cin_a = PN_FOLD * f;
                io.T_valid = (Ls + 3 + PN_FOLD - 1) / PN_FOLD;
                io.fold = PN_FOLD;
                io.fold_c = f;
                io.fold_o = 3;
                io.fold_T = Ls;
            } else {
                io.x_roff = padl[lvl];
                io.x2_pitch = pitchA[lvl];
                io.x2_roff = 1 + off;
                io.cin_a = f;
            }
            const bool tapped = r.stop_name && std::strcmp(r.stop_name, us_names[i]) == 0;
            if (i == 3 && !tapped) {  // last layer: + `out` 1x1 conv + softmax -> (B, 3, L0) fp32
                io.out_fmt = 2;
                io.y = y;
                io.y_ss = 3 * (int64_t)L0;
                io.y_cs = L0;
                io.head_w = m->pn_head_w;
                io.head_b = m->pn_head_b;
                if (r.go()) r.rc = tc_launch(tu, io, r.s);
                return r.rc;
            }
            if (run_layer(us_names[i], tu, io, Bu[i], Ls, 0, f, Ls)) return r.rc;
        }
        cur = Bu[i];
        cur_c = f;
        cur_len = Ls;
    }
    return r.rc;
}

static int run_pn(Runner &r, const float *x, float *y, Arena &ar) {
    if (r.precision != VP_PREC_FP32) return run_pn_tc(r, x, y, ar);
    vp_model *m = r.m;
    const int64_t B = r.B;
    const int L0 = m->in_samples;
    // SeisBench PhaseNet: manual pads before the stride-4 convs
    static const int padl[4] = {3, 2, 1, 2}, padr[4] = {3, 3, 3, 3};
    int len[5];
    len[0] = L0;
    for (int i = 0; i < 4; ++i) len[i + 1] = (len[i] + padl[i] + padr[i] - 7) / 4 + 1;
    float *b_inc = ar.take(B * 8 * len[0]);
    float *cat[4];  // cat[i]: (B, 2*C_i, len[i]) = [skip_i | up]
    float *dn[4];
    for (int i = 0; i < 4; ++i) {
        cat[i] = ar.take(B * 2 * kPnC[i] * len[i]);
        dn[i] = ar.take(B * kPnC[i] * len[i + 1]);
    }
    float *d4 = ar.take(B * 128 * len[4]);
    float *us[4];
    for (int i = 0; i < 4; ++i) us[i] = ar.take(B * kPnC[3 - i] * len[3 - i]);
    if (r.dry) return VP_OK;
    if (int rc = ar.check()) return rc;  // every buffer is laid out: refuse an undersized workspace BEFORE any launch

    r.conv(m->inc, 1, 1, 1, ACT_RELU, nullptr, nullptr, 0, x, 3 * (int64_t)L0, 0, L0, L0, 3, L0, L0, b_inc,
           8 * (int64_t)L0, 0, 1, 0, 0);
    if (r.tap("inc", b_inc, B * 8 * L0)) return r.rc;
    static const char *ds_names[5] = {"down0_same", "down1_same", "down2_same", "down3_same", "down4_same"};
    static const char *dd_names[4] = {"down0_down", "down1_down", "down2_down", "down3_down"};
    const float *cur = b_inc;
    int cur_c = 8;
    for (int i = 0; i < 5; ++i) {
        const int f = kPnC[i];
        if (i < 4) {
            // conv_same writes the skip half of the concat buffer (batch stride 2*f*len)
            r.conv(m->down_same[i], 1, 1, 1, ACT_RELU, nullptr, nullptr, 0, cur, (int64_t)cur_c * len[i], 0, len[i], len[i],
                   3, len[i], len[i], cat[i], 2 * (int64_t)f * len[i], 0, 1, 0, 0);
            if (r.tap(ds_names[i], cat[i], B * 2 * f * len[i])) return r.rc;  // NOTE: strided tap (skip half valid)
            r.conv(m->down_down[i], 4, 1, 1, ACT_RELU, nullptr, nullptr, 0, cat[i], 2 * (int64_t)f * len[i], 0, len[i],
                   len[i], padl[i], len[i + 1], len[i + 1], dn[i], (int64_t)f * len[i + 1], 0, 1, 0, 0);
            if (r.tap(dd_names[i], dn[i], B * f * len[i + 1])) return r.rc;
            cur = dn[i];
            cur_c = f;
        } else {
            r.conv(m->down_same[i], 1, 1, 1, ACT_RELU, nullptr, nullptr, 0, cur, (int64_t)cur_c * len[i], 0, len[i], len[i],
                   3, len[i], len[i], d4, (int64_t)f * len[i], 0, 1, 0, 0);
            if (r.tap(ds_names[i], d4, B * f * len[i])) return r.rc;
            cur = d4;
            cur_c = f;
        }
    }
    static const char *uc_names[4] = {"up0_cat", "up1_cat", "up2_cat", "up3_cat"};
    static const char *us_names[4] = {"up0_same", "up1_same", "up2_same", "up3_same"};
    int cur_len = len[4];
    for (int i = 0; i < 4; ++i) {
        const int lvl = 3 - i;
        const int f = kPnC[lvl];
        const int Ls = len[lvl];
        const int Lt = (cur_len - 1) * 4 + 7 - 3;  // ConvTranspose1d output, cropped [1:-2]
        const int off = (Lt - Ls) / 2;
        if (off < 0 || off + Ls > Lt) {
            set_error("PhaseNet: skip merge mismatch (Lt=%d, Ls=%d)", Lt, Ls);
            return VP_ERR_ARG;
        }
        if (r.go()) {
            ConvTP p;
            p.x = cur;
            p.x_bs = (int64_t)cur_c * cur_len;
            p.w = r.W(m->up_t[i].w);
            p.bias = r.W(m->up_t[i].b);
            p.y = cat[lvl] + (int64_t)f * Ls;  // second half of the concat buffer
            p.y_bs = 2 * (int64_t)f * Ls;
            p.Lin = cur_len;
            p.Lout = Ls;
            p.shift = 1 + off;
            r.rc = launch_convt_fp32(cur_c, f, p, (int)B, r.s);
        }
        if (r.tap(uc_names[i], cat[lvl], B * 2 * f * Ls)) return r.rc;
        r.conv(m->up_same[i], 1, 1, 1, ACT_RELU, nullptr, nullptr, 0, cat[lvl], 2 * (int64_t)f * Ls, 0, Ls, Ls, 3, Ls, Ls,
               us[i], (int64_t)f * Ls, 0, 1, 0, 0);
        if (r.tap(us_names[i], us[i], B * f * Ls)) return r.rc;
        cur = us[i];
        cur_c = f;
        cur_len = Ls;
    }
    r.conv(m->outc, 1, 1, 1, ACT_SOFTMAX3, nullptr, nullptr, 0, cur, 8 * (int64_t)L0, 0, L0, L0, 0, L0, L0, y,
           3 * (int64_t)L0, 0, 1, 0, 0);
    return r.rc;
}

static int run_forward(vp_model *m, const float *x, int64_t B, float *y, void *ws, int64_t ws_bytes, int precision,
                       const char *stop, Tap *hit, cudaStream_t s, int64_t *need_bytes, int keep_lo = 0, int keep_hi = 1 << 30,
                       const Runner *src = nullptr) {
    if (precision != VP_PREC_FP32 && precision != VP_PREC_F16X3 && precision != VP_PREC_BF16) {
        set_error("unknown precision mode %d", precision);
        return VP_ERR_ARG;
    }
    Runner r;
    r.precision = precision;
    r.keep_lo = keep_lo;
    r.keep_hi = keep_hi;
    if (src) {
        r.trace = src->trace;
        r.trace_dtype = src->trace_dtype;
        r.peak_scope = src->peak_scope;
        r.taper = src->taper;
        r.ch_stride = src->ch_stride;
        r.starts = src->starts;
    }
    r.m = m;
    r.s = s;
    r.B = (int)B;
    r.dry = need_bytes != nullptr;
    r.stop_name = stop;
    Arena ar{(char *)ws, 0, ws_bytes};
    if (r.dry) ar.base = nullptr;
    int rc = (m->kind == VP_KIND_EQTRANSFORMER) ? run_eqt(r, x, y, ar) : run_pn(r, x, y, ar);
    if (need_bytes) *need_bytes = align_up(ar.used, 256);
    if (rc != VP_OK) return rc;
    if (!r.dry && ar.used > ws_bytes) {
        set_error("forward workspace too small: need %lld bytes, have %lld", (long long)ar.used, (long long)ws_bytes);
        return VP_ERR_WORKSPACE;
    }
    if (hit) *hit = r.hit;
    return VP_OK;
}

// ============================================================================================ C ABI
extern "C" int vp_version(void) { return 100; }
extern "C" const char *vp_last_error(void) { return g_error.c_str(); }
extern "C" int64_t vp_launch_count(int reset) {
    const int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

extern "C" int vp_kernel_timing(int enable) {
    for (auto &sp : vp::g_kspans) {
        if (sp.e0) vp::g_kevent_pool.push_back(sp.e0);
        if (sp.e1) vp::g_kevent_pool.push_back(sp.e1);
    }
    vp::g_kspans.clear();
    vp::g_ktimer_on = enable != 0;
    return VP_OK;
}

extern "C" const char *vp_kernel_class_names(void) { return vp::KCLASS_NAMES; }

extern "C" int vp_kernel_timing_read(int kclass, double *total_ms, int64_t *launches) {
    VP_REQUIRE(kclass >= 0 && kclass < vp::KC_COUNT && total_ms && launches, VP_ERR_ARG, "vp_kernel_timing_read: bad argument");
    double tot = 0.0;
    int64_t cnt = 0;
    for (auto &sp : vp::g_kspans) {
        if (sp.cls != kclass || !sp.e1) continue;
        VP_CUDA_CHECK(cudaEventSynchronize(sp.e1));
        float ms = 0.f;
        VP_CUDA_CHECK(cudaEventElapsedTime(&ms, sp.e0, sp.e1));
        tot += ms;
        ++cnt;
    }
    *total_ms = tot;
    *launches = cnt;
    return VP_OK;
}
extern "C" int64_t vp_model_expected_floats(int kind) {
    return kind == VP_KIND_EQTRANSFORMER ? EQT_FLOATS : kind == VP_KIND_PHASENET ? PN_FLOATS : -1;
}

extern "C" int vp_model_create(int kind, const float *weights, int64_t n_floats, int device, vp_model **out) {
    VP_REQUIRE(out && weights, VP_ERR_ARG, "vp_model_create: null pointer");
    VP_REQUIRE(kind == VP_KIND_EQTRANSFORMER || kind == VP_KIND_PHASENET, VP_ERR_ARG, "vp_model_create: unknown kind %d", kind);
    VP_REQUIRE(n_floats == vp_model_expected_floats(kind), VP_ERR_ARG,
               "vp_model_create: expected %lld weight floats for kind %d, got %lld",
               (long long)vp_model_expected_floats(kind), kind, (long long)n_floats);
    int ndev = 0;
    VP_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    VP_REQUIRE(device >= 0 && device < ndev, VP_ERR_CUDA, "vp_model_create: CUDA device %d not available (%d devices)", device, ndev);
    struct DeviceGuard {  // the caller's current device is restored on every path out of this function
        int prev = -1;
        ~DeviceGuard() {
            if (prev >= 0) cudaSetDevice(prev);
        }
    } guard;
    VP_CUDA_CHECK(cudaGetDevice(&guard.prev));
    VP_CUDA_CHECK(cudaSetDevice(device));
    vp_model *m = new vp_model();
    m->kind = kind;
    m->device = device;
    m->in_samples = kind == VP_KIND_EQTRANSFORMER ? 6000 : 3001;
    Cursor cur{weights, n_floats};
    Packed pk;
    if (kind == VP_KIND_EQTRANSFORMER) {
        int rc = build_eqt(m, cur, pk);
        if (rc != VP_OK) {
            delete m;
            return rc;
        }
    } else {
        int rc = build_pn(m, cur, pk);
        if (rc != VP_OK) {
            delete m;
            return rc;
        }
    }
    if (cur.left != 0) {
        set_error("vp_model_create: weight walk left %lld floats (layout mismatch)", (long long)cur.left);
        delete m;
        return VP_ERR_ARG;
    }
    m->n_packed = (int64_t)pk.host.size();
    cudaError_t e = cudaMalloc(&m->d_weights, (pk.host.size() + 64) * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(m->d_weights, pk.host.data(), pk.host.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("vp_model_create: uploading weights failed: %s", cudaGetErrorString(e));
        if (m->d_weights) cudaFree(m->d_weights);
        delete m;
        return VP_ERR_CUDA;
    }
    if (kind == VP_KIND_EQTRANSFORMER) {
        for (int set = 0; set < 2; ++set) {
            int rc = upload_tc(m->tc[set]);
            if (rc == VP_OK) rc = decb_upload(m->tc[set].decb);
            if (rc == VP_OK) rc = decb2_upload(m->tc[set].decb2);
            if (rc == VP_OK) rc = enca_upload(m->tc[set].enca);
            if (rc == VP_OK) rc = deca_upload(m->tc[set].deca);
            if (rc != VP_OK) {
                vp_model_destroy(m);
                return rc;
            }
        }
    }
    if (kind == VP_KIND_PHASENET) {
        for (int set = 0; set < 2; ++set) {
            int rc = upload_pn_tc(m->pn_tc[set]);
            if (rc != VP_OK) {
                vp_model_destroy(m);
                return rc;
            }
        }
    }
    *out = m;
    return VP_OK;
}

extern "C" int vp_model_destroy(vp_model *m) {
    if (!m) return VP_OK;
    if (m->d_weights) cudaFree(m->d_weights);
    for (int set = 0; set < 2; ++set) {
        if (m->tc[set].d_w) cudaFree(m->tc[set].d_w);
        if (m->tc[set].d_b) cudaFree(m->tc[set].d_b);
        if (m->tc[set].d_res_par) cudaFree(m->tc[set].d_res_par);
        decb_free(m->tc[set].decb);
        decb2_free(m->tc[set].decb2);
        enca_free(m->tc[set].enca);
        deca_free(m->tc[set].deca);
        if (m->pn_tc[set].d_w) cudaFree(m->pn_tc[set].d_w);
        if (m->pn_tc[set].d_b) cudaFree(m->pn_tc[set].d_b);
    }
    delete m;
    return VP_OK;
}
extern "C" int vp_model_kind(const vp_model *m) { return m ? m->kind : VP_ERR_ARG; }
extern "C" int vp_model_in_samples(const vp_model *m) { return m ? m->in_samples : VP_ERR_ARG; }
extern "C" const char *vp_forward_tap_names(const vp_model *m) { return m ? m->tap_names.c_str() : ""; }

extern "C" int64_t vp_forward_workspace_bytes(const vp_model *m, int64_t n_windows, int precision) {
    if (!m || n_windows < 0) return VP_ERR_ARG;
    const int64_t B = std::min<int64_t>(std::max<int64_t>(n_windows, 1), MAX_CHUNK);
    int64_t need = 0;
    int rc = run_forward(const_cast<vp_model *>(m), nullptr, B, nullptr, nullptr, 0, precision, nullptr, nullptr, 0, &need);
    return rc == VP_OK ? need : rc;
}

extern "C" int vp_forward(vp_model *m, const float *x, int64_t n_windows, float *y, void *workspace,
                          int64_t workspace_bytes, int precision, void *stream) {
    VP_REQUIRE(m && x && y && workspace, VP_ERR_ARG, "vp_forward: null pointer");
    const int64_t L = m->in_samples;
    for (int64_t b0 = 0; b0 < n_windows; b0 += MAX_CHUNK) {
        const int64_t nb = std::min<int64_t>(MAX_CHUNK, n_windows - b0);
        int rc = run_forward(m, x + b0 * 3 * L, nb, y + b0 * 3 * L, workspace, workspace_bytes, precision, nullptr,
                             nullptr, (cudaStream_t)stream, nullptr);
        if (rc != VP_OK) return rc;
    }
    return VP_OK;
}

extern "C" int vp_forward_range(vp_model *m, const float *x, int64_t n_windows, float *y, void *workspace,
                                int64_t workspace_bytes, int precision, int64_t keep_lo, int64_t keep_hi, void *stream) {
    VP_REQUIRE(m && x && y && workspace, VP_ERR_ARG, "vp_forward_range: null pointer");
    const int64_t L = m->in_samples;
    VP_REQUIRE(keep_lo >= 0 && keep_lo <= keep_hi, VP_ERR_ARG, "vp_forward_range: bad sample range [%lld, %lld)", (long long)keep_lo,
               (long long)keep_hi);
    for (int64_t b0 = 0; b0 < n_windows; b0 += MAX_CHUNK) {
        const int64_t nb = std::min<int64_t>(MAX_CHUNK, n_windows - b0);
        int rc = run_forward(m, x + b0 * 3 * L, nb, y + b0 * 3 * L, workspace, workspace_bytes, precision, nullptr, nullptr,
                             (cudaStream_t)stream, nullptr, (int)std::min<int64_t>(keep_lo, L), (int)std::min<int64_t>(keep_hi, L));
        if (rc != VP_OK) return rc;
    }
    return VP_OK;
}

extern "C" int vp_slice_forward(vp_model *m, const void *trace, int dtype, int64_t n_samples, int64_t ch_stride,
                                const int64_t *starts, int64_t n_windows, int peak_scope, int taper, float *y, void *workspace,
                                int64_t workspace_bytes, int precision, int64_t keep_lo, int64_t keep_hi, void *stream) {
    VP_REQUIRE(m && trace && starts && y && workspace, VP_ERR_ARG, "vp_slice_forward: null pointer");
    VP_REQUIRE(precision == VP_PREC_F16X3 || precision == VP_PREC_BF16, VP_ERR_UNSUPPORTED,
               "vp_slice_forward: implemented for the tensor-core modes (use vp_slice_normalize + vp_forward otherwise)");
    VP_REQUIRE(dtype == VP_DTYPE_F32 || dtype == VP_DTYPE_I32, VP_ERR_ARG, "vp_slice_forward: unknown dtype %d", dtype);
    VP_REQUIRE(keep_lo >= 0 && keep_lo <= keep_hi, VP_ERR_ARG, "vp_slice_forward: bad sample range [%lld, %lld)", (long long)keep_lo,
               (long long)keep_hi);
    (void)n_samples;
    const int64_t L = m->in_samples;
    for (int64_t b0 = 0; b0 < n_windows; b0 += MAX_CHUNK) {
        const int64_t nb = std::min<int64_t>(MAX_CHUNK, n_windows - b0);
        Runner src;
        src.trace = trace;
        src.trace_dtype = dtype;
        src.peak_scope = peak_scope;
        src.taper = taper;
        src.ch_stride = ch_stride;
        src.starts = starts + b0;
        int rc = run_forward(m, nullptr, nb, y + b0 * 3 * L, workspace, workspace_bytes, precision, nullptr, nullptr, (cudaStream_t)stream,
                             nullptr, (int)std::min<int64_t>(keep_lo, L), (int)std::min<int64_t>(keep_hi, L), &src);
        if (rc != VP_OK) return rc;
    }
    return VP_OK;
}

extern "C" int vp_forward_tap(vp_model *m, const float *x, int64_t n_windows, float *y, void *workspace,
                              int64_t workspace_bytes, int precision, const char *tap_name, float *tap_out,
                              int64_t tap_capacity, int64_t *tap_floats, void *stream) {
    VP_REQUIRE(m && x && y && workspace && tap_name && tap_out && tap_floats, VP_ERR_ARG, "vp_forward_tap: null pointer");
    VP_REQUIRE(n_windows <= MAX_CHUNK, VP_ERR_ARG, "vp_forward_tap: at most %lld windows", (long long)MAX_CHUNK);
    Tap hit{nullptr, nullptr, 0};
    int rc = run_forward(m, x, n_windows, y, workspace, workspace_bytes, precision, tap_name, &hit, (cudaStream_t)stream, nullptr);
    if (rc != VP_OK) return rc;
    VP_REQUIRE(hit.ptr != nullptr, VP_ERR_ARG, "vp_forward_tap: unknown tap '%s' (have: %s)", tap_name, m->tap_names.c_str());
    VP_REQUIRE(hit.floats <= tap_capacity, VP_ERR_CAPACITY, "vp_forward_tap: tap '%s' needs %lld floats, capacity %lld",
               tap_name, (long long)hit.floats, (long long)tap_capacity);
    VP_CUDA_CHECK(cudaMemcpyAsync(tap_out, hit.ptr, hit.floats * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    *tap_floats = hit.floats;
    return VP_OK;
}
