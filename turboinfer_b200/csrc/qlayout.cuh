// qlayout.cuh -- the packed INT4/INT8 weight layout in HBM, shared by the pack / unpack kernels and the
// streaming GEMV.  This replaces the reference's "one int32 (INT4) / one int8 per element, [K][N] row-major"
// tensors (src/optimize/quantization.cpp:45-46) with a layout made for one thing: every SM streams ONE
// contiguous slab of the matrix with 128-bit accesses, in exactly the order its warps consume it.
//
// Terms
//   item        one output column x one "superchunk" of k: 32 lanes x 16 B = 512 B
//               INT4: 1024 k per item (32 nibbles per lane), INT8: 512 k per item (16 bytes per lane)
//   lane vector lane l of an item holds k = s*KSC + 128*j + 4*l + e  (j < KSC/128, e < 4), so that the matching
//               activations are the float4 at x[s*KSC + 128*j + 4*l] -- consecutive lanes, no bank conflicts
//               INT4: word j/2, nibble (j%2)*4 + e, stored as u = q + 8;  INT8: word j, byte e, stored as q + 128
//   CTA slab    the matrix is cut by columns into P slabs (units of 4 columns, P = #SMs when N allows);
//               slab p is one contiguous byte range.  Inside it, items are numbered s-major (i = s*ncols + c),
//               dealt to the 16 consumer warps as contiguous ranges, and laid out round by round:
//               [round][warp][<=4 items][lane][16 B], a round being what the 16 warps consume from one
//               pipeline stage (<= 32 KiB, one bulk async copy).
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
    int ksc;      // k per superchunk: 1024 (INT4) / 512 (INT8)
    int nsc;      // superchunks = ceil(K / ksc)
    int U;        // column units of 4 = ceil(N / 4)
    int P;        // slabs (= CTAs of the GEMV)
};

TIB_HD QLayout make_layout(int K, int N, int bits, int num_sms) {
    QLayout L;
    L.K = K; L.N = N; L.bits = bits;
    L.ksc = bits == 4 ? 1024 : 512;
    L.nsc = (K + L.ksc - 1) / L.ksc;
    L.U = (N + 3) / 4;
    L.P = L.U < num_sms ? L.U : num_sms;
    return L;
}

TIB_HD size_t layout_bytes(const QLayout& L) { return (size_t)4 * L.U * L.nsc * kItemBytes; }
TIB_HD int layout_kpad(const QLayout& L) { return L.nsc * L.ksc; }

struct Slab {
    int col0;     // first column
    int ncols;    // columns in this slab (multiple of 4, may run past N: padding columns are zero)
    int T;        // items = ncols * nsc
    int b, m;     // T = 16*b + m: warps w < m own b+1 items, the others b
    int rounds;   // pipeline stages this slab is streamed in
    size_t byte0; // offset of the slab in the packed buffer
};

TIB_HD Slab make_slab(const QLayout& L, int p) {
    Slab s;
    int u0 = (int)((long long)p * L.U / L.P);
    int u1 = (int)((long long)(p + 1) * L.U / L.P);
    s.col0 = 4 * u0;
    s.ncols = 4 * (u1 - u0);
    s.T = s.ncols * L.nsc;
    s.b = s.T / kConsumerWarps;
    s.m = s.T % kConsumerWarps;
    int nmax = s.b + (s.m > 0 ? 1 : 0);
    s.rounds = (nmax + kItemsPerRound - 1) / kItemsPerRound;
    s.byte0 = (size_t)4 * u0 * L.nsc * kItemBytes;
    return s;
}

TIB_HD int slab_max_items(const QLayout& L) {  // upper bound of Slab::T over all slabs
    int umax = (L.U + L.P - 1) / L.P;
    return 4 * umax * L.nsc;
}

// items warp `w` consumes in round `r`, given it owns n items in total
TIB_HD int round_items(int n, int r) {
    int g = n - kItemsPerRound * r;
    return g < 0 ? 0 : (g > kItemsPerRound ? kItemsPerRound : g);
}
TIB_HD int warp_items(const Slab& s, int w) { return s.b + (w < s.m ? 1 : 0); }
TIB_HD int warp_first_item(const Slab& s, int w) { return w * s.b + (w < s.m ? w : s.m); }
// items in round r over all warps, and the item offset of warp w inside round r
TIB_HD int round_total(const Slab& s, int r) {
    return s.m * round_items(s.b + 1, r) + (kConsumerWarps - s.m) * round_items(s.b, r);
}
TIB_HD int round_warp_offset(const Slab& s, int r, int w) {
    int lo = w < s.m ? w : s.m;
    int hi = w - lo;
    return lo * round_items(s.b + 1, r) + hi * round_items(s.b, r);
}

// INT4: power-of-16 position p of element (j, e) inside its 32-bit word after the w / (w >> 12) split:
// nibbles 0..4 are read in place (p = 0..4), nibbles 5..7 from w >> 12 at p = 2..4.
TIB_HD int q4_pos(int j, int e) {
    int nib = (j & 1) * 4 + e;
    return nib <= 4 ? nib : nib - 3;
}

}  // namespace tib
