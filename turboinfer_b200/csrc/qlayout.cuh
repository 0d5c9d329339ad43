// qlayout.cuh -- the packed INT4/INT8 weight layout in HBM, shared by the pack / unpack kernels and the
// streaming GEMV.  This replaces the reference's "one int32 (INT4) / one int8 per element, [K][N] row-major"
// tensors (src/optimize/quantization.cpp:45-46) with a layout made for one thing: every SM streams ONE
// contiguous slab of the matrix with 128-bit accesses, in exactly the order its warps consume it.
//
// Terms
//   unit        4 consecutive output columns
//   chunk       kc consecutive k: 256 (INT4) / 128 (INT8)
//   item        one unit x one chunk = 32 lanes x 16 B = 512 B.  Lane l = (c = l & 3, s = l >> 2) holds, for column
//               4*unit + c, the kc/8 values k = chunk*kc + s*(kc/8) + [0, kc/8):
//               INT4: 32 nibbles u = q + off; word j (0..3), nibble 2i   -> k_local = 8j + i       (i < 4)
//                                                        nibble 2i+1 -> k_local = 8j + 4 + i
//                     so (w & 0x0F0F0F0F) and ((w >> 4) & 0x0F0F0F0F) are two dp4a operands of 4 consecutive k each
//               INT8: 16 bytes q (two's complement); word j (0..3), byte i -> k_local = 4j + i
//   group       4 consecutive units of a slab (16 columns); the last group of a slab may have 1..3 units
//   quad        the <= 4 items of one group at one chunk: they share their activations, which a warp therefore loads
//               from shared memory once per quad
//   CTA slab    the matrix is cut by columns into P slabs (whole units, P = #SMs when N allows); slab p is one
//               contiguous byte range.  Inside it, quads are numbered group-major (q = group*nchunks + chunk), dealt
//               to the 16 consumer warps as contiguous ranges, and laid out round by round:
//               [round][warp][<= 4 items of the warp's quad][lane][16 B], a round being what the 16 warps consume
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
    const int u0 = (int)((long long)p * L.U / L.P);
    const int u1 = (int)((long long)(p + 1) * L.U / L.P);
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

// position of element k_local (0 .. kc/8) of a lane inside its 16 bytes: word index and bit shift
TIB_HD void lane_elem_pos(int bits, int k_local, int& word, int& shift) {
    if (bits == 4) {
        word = k_local >> 3;
        const int r = k_local & 7;
        shift = r < 4 ? 8 * r : 8 * (r - 4) + 4;
    } else {
        word = k_local >> 2;
        shift = 8 * (k_local & 3);
    }
}

// The activation vector is staged in shared memory as three 8-bit digit planes of a 24-bit fixed-point value
// (gemv.cuh).  Byte address of the 32-bit word holding digit d of the four values k .. k+3 (k % 4 == 0), laid out so
// that the 8 k-slices of a chunk are contiguous 16-byte vectors (conflict-free LDS.128 from lanes (c, s)):
//   INT4: [chunk][digit][half h][slice s][word j]   k = chunk*256 + s*32 + j*8 + h*4
//   INT8: [chunk][digit][slice s][word j]           k = chunk*128 + s*16 + j*4
TIB_HD int xdigit_word_offset(int bits, int k, int d) {
    if (bits == 4) {
        const int chunk = k >> 8, s = (k >> 5) & 7, j = (k >> 3) & 3, h = (k >> 2) & 1;
        return ((((chunk * 3 + d) * 2 + h) * 8 + s) * 4 + j) * 4;
    }
    const int chunk = k >> 7, s = (k >> 4) & 7, j = (k >> 2) & 3;
    return (((chunk * 3 + d) * 8 + s) * 4 + j) * 4;
}
TIB_HD size_t xdigit_bytes(const QLayout& L) { return (size_t)3 * layout_kpad(L); }

}  // namespace tib
