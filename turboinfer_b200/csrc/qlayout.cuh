// qlayout.cuh -- the packed INT4/INT8 weight layout in HBM, shared by the pack / unpack kernels and the
// streaming GEMV.  This replaces the reference's "one int32 (INT4) / one int8 per element, [K][N] row-major"
// tensors (src/optimize/quantization.cpp:45-46) with a layout made for one thing: every SM streams ONE
// contiguous slab of the matrix with 128-bit accesses, in exactly the order its warps consume it.
//
// Terms
//   unit        4 consecutive output columns
//   chunk       kc consecutive k: 256 (INT4) / 128 (INT8)
//   group       4 consecutive units of a slab (16 columns); the last group of a slab may have nl = 1..3 units
//   quad        one group x one chunk: nl * 512 B, stored as 4 k-items in warp-MMA fragment order (see kitem_word_elem
//               below): a k-item is the A operand of mma.sync.m16n8k32 (rows = the group's columns), INT4 two of them
//               (low / high nibbles).  "item" below is the byte-accounting unit of 512 B: a quad counts nl items.
//   CTA slab    the matrix is cut by columns into P slabs (whole units, P = #SMs when N allows); slab p is one
//               contiguous byte range.  Inside it, quads are numbered group-major (q = group*nchunks + chunk), dealt
//               to the 16 consumer warps as contiguous ranges, and laid out round by round:
//               [round][warp][the warp's quad: 4 k-items], a round being what the 16 warps consume
//               from one pipeline stage (<= 32 KiB, one bulk async copy).  No padding items are stored.
// K is padded to whole chunks with zeros (q = 0); every Llama-family K is a multiple of 256, so no padding bytes
// are streamed for the benchmark shapes.
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
#define TIB_HD __host__ __device__ __forceinline__
#else
#define TIB_HD inline
#endif

namespace tib {

constexpr int kConsumerWarps = 16;
constexpr int kItemsPerRound = 4;     // per warp
constexpr int kItemBytes = 512;       // 32 lanes x 16 B
constexpr int kStageBytes = kConsumerWarps * kItemsPerRound * kItemBytes;  // 32 KiB

struct QLayout {
    int K, N;     // logical sizes (y[N] = x[K] . W[K][N])
    int bits;     // 4 or 8
    int kc;       // k per chunk: 256 (INT4) / 128 (INT8)
    int nchunks;  // chunks = ceil(K / kc)
    int U;        // column units of 4 = ceil(N / 4)
    int P;        // slabs (= CTAs of the GEMV)
};

TIB_HD QLayout make_layout(int K, int N, int bits, int num_sms) {
    QLayout L;
    L.K = K; L.N = N; L.bits = bits;
    L.kc = bits == 4 ? 256 : 128;
    L.nchunks = (K + L.kc - 1) / L.kc;
    L.U = (N + 3) / 4;
    L.P = L.U < num_sms ? L.U : num_sms;
    return L;
}

TIB_HD size_t layout_bytes(const QLayout& L) { return (size_t)L.U * L.nchunks * kItemBytes; }
TIB_HD int layout_kpad(const QLayout& L) { return L.nchunks * L.kc; }

struct Slab {
    int unit0;    // first unit
    int nunits;   // units in this slab
    int col0;     // = 4 * unit0
    int ncols;    // = 4 * nunits (may run past N: padding columns are zero)
    int ngroups;  // groups of 4 units = ceil(nunits / 4)
    int nlast;    // units in the last group (1..4)
    int qfull;    // quads with index >= qfull belong to the last group
    int Tq;       // quads = ngroups * nchunks
    int bq, mq;   // Tq = 16*bq + mq: warps w < mq own bq+1 quads, the others bq
    int rounds;   // pipeline stages this slab is streamed in (one quad per warp per round)
    size_t byte0; // offset of the slab in the packed buffer
};

TIB_HD Slab make_slab(const QLayout& L, int p) {
    Slab s;
    // 32-bit arithmetic: p * U < 2^31 for every supported shape (U < 2^22 units, p < 512)
    const int u0 = (int)((unsigned)p * (unsigned)L.U / (unsigned)L.P);
    const int u1 = (int)((unsigned)(p + 1) * (unsigned)L.U / (unsigned)L.P);
    s.unit0 = u0;
    s.nunits = u1 - u0;
    s.col0 = 4 * u0;
    s.ncols = 4 * (u1 - u0);
    s.ngroups = (s.nunits + 3) / 4;
    s.nlast = s.nunits - 4 * (s.ngroups - 1);
    s.qfull = (s.ngroups - 1) * L.nchunks;
    s.Tq = s.ngroups * L.nchunks;
    s.bq = s.Tq / kConsumerWarps;
    s.mq = s.Tq % kConsumerWarps;
    s.rounds = s.bq + (s.mq > 0 ? 1 : 0);
    s.byte0 = (size_t)u0 * L.nchunks * kItemBytes;
    return s;
}

TIB_HD int slab_max_units(const QLayout& L) { return (L.U + L.P - 1) / L.P; }

TIB_HD int warp_quads(const Slab& s, int w) { return s.bq + (w < s.mq ? 1 : 0); }
TIB_HD int warp_first_quad(const Slab& s, int w) { return w * s.bq + (w < s.mq ? w : s.mq); }
// items of the quad warp w handles in round r (0 when it has none)
TIB_HD int round_warp_items(const Slab& s, int r, int w) {
    if (r >= warp_quads(s, w)) return 0;
    return warp_first_quad(s, w) + r >= s.qfull ? s.nlast : 4;
}
// item offset of warp w inside stage r (w = 16: items in the stage)
TIB_HD int round_warp_offset(const Slab& s, int r, int w) {
    int off = 0;
    for (int i = 0; i < w; ++i) off += round_warp_items(s, r, i);
    return off;
}
TIB_HD int round_total(const Slab& s, int r) { return round_warp_offset(s, r, kConsumerWarps); }

// ---- inside a quad: the warp-MMA fragment order ---------------------------------------------------------------
// A quad (16 columns x one chunk; the slab's last group may have nl = 1..3 units = 4*nl columns) is stored as 4
// consecutive k-items.  A k-item covers 64 k (INT4) / 32 k (INT8) of the group's columns and is exactly the A operand
// of mma.sync.m16n8k32 (rows = the group's columns, 32 k per instruction; INT4: the low nibbles of a word are the
// fragment of the first 32 k, the high nibbles that of the next 32 k):
//   lane = (g = lane >> 2, t = lane & 3); fragment register j (0..3), byte b (0..3):
//       row (column in the group) = g + 8 * (j & 1),   k in the 32-k block = 4*t + b + 16 * (j >> 1)
//   full group (nl = 4):  512 B, lane l holds registers j = 0..3 at bytes [16 l, 16 l + 16)          (one LDS.128)
//   ragged group:         rows 0..7 first ("part A": registers j = 0, 2 of lanes g < min(8, 4 nl), 8 B per lane), then
//                         rows 8..11 for nl = 3 ("part B": registers j = 1, 3 of lanes g < 4) -- nl * 128 B, no padding
// so a quad occupies nl * 512 B whatever nl is, as many bytes as its elements need.
TIB_HD int kitem_bytes(int nl) { return nl * 128; }
TIB_HD int kitem_k(int bits) { return bits == 4 ? 64 : 32; }

// Inverse map used by the pack / unpack kernels: 32-bit word `widx` of a k-item of a group with nl units, element i of
// the word (INT4: nibble i, bits [4i, 4i+4); INT8: byte i) -> row (column in the group) and k offset inside the k-item.
TIB_HD void kitem_word_elem(int bits, int nl, int widx, int i, int& row, int& kk) {
    int lane, j;
    if (nl == 4) { lane = widx >> 2; j = widx & 3; }
    else {
        const int na = nl == 1 ? 32 : 64;           // words of part A
        if (widx < na) { lane = widx >> 1; j = 2 * (widx & 1); }
        else { lane = (widx - na) >> 1; j = 2 * ((widx - na) & 1) + 1; }
    }
    const int g = lane >> 2, t = lane & 3;
    row = g + 8 * (j & 1);
    if (bits == 4) kk = 4 * t + (i >> 1) + 16 * (j >> 1) + 32 * (i & 1);
    else kk = 4 * t + i + 16 * (j >> 1);
}

#ifdef __CUDACC__
// Walks the 32-bit words of the quads that consumer warp `warp` of slab `slab` owns (the pack / unpack kernels mirror
// the GEMV's 16 consumer warps): f(byte offset of the word in the packed buffer, nl, group, chunk, k-item, word in k-item)
template <typename F>
__device__ __forceinline__ void for_each_quad_word(const QLayout& L, const Slab& slab, int warp, int lane, F&& f) {
    const int nq = warp_quads(slab, warp), fq = warp_first_quad(slab, warp);
    size_t stage_base = 0;    // byte offset of round r inside the slab
    for (int r = 0; r < nq; ++r) {
        const int q = fq + r;
        const int grp = q / L.nchunks, chunk = q - grp * L.nchunks;
        const int nl = q >= slab.qfull ? slab.nlast : 4;
        const size_t qoff = slab.byte0 + stage_base + (size_t)round_warp_offset(slab, r, warp) * kItemBytes;
        stage_base += (size_t)round_total(slab, r) * kItemBytes;
        const int wpk = nl * 32;   // words per k-item
        for (int wq = lane; wq < 4 * wpk; wq += 32) f(qoff + (size_t)wq * 4, nl, grp, chunk, wq / wpk, wq % wpk);
    }
}
#endif

// The activation vector is staged in shared memory as three SIGNED 8-bit digit planes of a 24-bit fixed-point value,
// xf = d2*65536 + d1*256 + d0 with every digit in [-128, 127] (gemv.cuh), in the order the B fragments of the MMAs want:
// lane (g, t) multiplies digit plane g (columns g >= 3 of the product are unused) and needs, per 32-k block, the words
// k = 4t .. 4t+3 and k = 16+4t .. 16+4t+3.
//   INT4: [k / 64][digit][t][4 words: k%64 = 4t, 16+4t, 32+4t, 48+4t]     one LDS.128 per k-item (two MMAs)
//   INT8: [k / 32][digit][t][2 words: k%32 = 4t, 16+4t]                   one LDS.64 per k-item (one MMA)
// Byte address of the word holding digit d of the four values k .. k+3 (k % 4 == 0):
TIB_HD int xdigit_word_offset(int bits, int k, int d) {
    const int t = (k >> 2) & 3, q = k >> 4;
    if (bits == 4) return (((q >> 2) * 3 + d) * 4 + t) * 16 + (q & 3) * 4;
    return (((q >> 1) * 3 + d) * 4 + t) * 8 + (q & 1) * 4;
}
TIB_HD size_t xdigit_bytes(const QLayout& L) { return (size_t)3 * layout_kpad(L); }

}  // namespace tib
