// Host-side check of the packed weight layout (turboinfer_b200/csrc/qlayout.cuh), no GPU needed: the slab partition
// covers every unit once, the per-stage bookkeeping of producer and consumers agrees with the slab's byte size, and the
// fragment-order map inside a quad is a bijection onto (row, k) for full and ragged groups.
#include <cstdio>
#include <set>
#include <utility>
#include <vector>
#include "../../turboinfer_b200/csrc/qlayout.cuh"
using namespace tib;

static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++g_fail; } } while (0)

static void check_matrix(int K, int N, int bits, int sms) {
    const QLayout L = make_layout(K, N, bits, sms);
    CHECK(L.nchunks * L.kc >= K && (L.nchunks - 1) * L.kc < K);
    int next_unit = 0;
    size_t bytes = 0;
    for (int p = 0; p < L.P; ++p) {
        const Slab s = make_slab(L, p);
        CHECK(s.unit0 == next_unit && s.nunits >= 1);
        CHECK(s.byte0 == bytes);
        next_unit += s.nunits;
        CHECK(s.nlast >= 1 && s.nlast <= 4 && 4 * (s.ngroups - 1) + s.nlast == s.nunits);
        // quads dealt to the 16 warps: contiguous, complete
        int q = 0;
        for (int w = 0; w < kConsumerWarps; ++w) { CHECK(warp_first_quad(s, w) == q); q += warp_quads(s, w); }
        CHECK(q == s.Tq);
        // stages: every round fits a stage, the rounds add up to the slab
        size_t slab_bytes = 0;
        for (int r = 0; r < s.rounds; ++r) {
            const int items = round_total(s, r);
            CHECK(items >= 1 && items * kItemBytes <= kStageBytes);
            int off = 0;
            for (int w = 0; w < kConsumerWarps; ++w) { CHECK(round_warp_offset(s, r, w) == off); off += round_warp_items(s, r, w); }
            slab_bytes += (size_t)items * kItemBytes;
        }
        CHECK(slab_bytes == (size_t)s.nunits * L.nchunks * kItemBytes);
        bytes += slab_bytes;
    }
    CHECK(next_unit == L.U);
    CHECK(bytes == layout_bytes(L));
    // the activation digits: every (k, digit) word has its own place inside 3 * kpad bytes
    std::set<int> seen;
    for (int k = 0; k < layout_kpad(L); k += 4)
        for (int d = 0; d < 3; ++d) {
            const int off = xdigit_word_offset(bits, k, d);
            CHECK(off >= 0 && off + 4 <= (int)xdigit_bytes(L) && off % 4 == 0);
            CHECK(seen.insert(off).second);
        }
}

static void check_kitem(int bits) {
    const int nel = bits == 4 ? 8 : 4, kik = kitem_k(bits);
    for (int nl = 1; nl <= 4; ++nl) {
        std::set<std::pair<int, int>> seen;
        const int words = kitem_bytes(nl) / 4;
        for (int w = 0; w < words; ++w)
            for (int i = 0; i < nel; ++i) {
                int row, kk;
                kitem_word_elem(bits, nl, w, i, row, kk);
                CHECK(row >= 0 && row < 4 * nl && kk >= 0 && kk < kik);
                CHECK(seen.insert({row, kk}).second);
            }
        CHECK((int)seen.size() == 4 * nl * kik);   // exactly the elements of the k-item: no padding bytes
    }
}

int main() {
    for (int bits : {4, 8}) {
        check_kitem(bits);
        for (int sms : {148, 132, 7})
            for (auto kn : std::vector<std::pair<int, int>>{{4096, 4096}, {4096, 12288}, {4096, 22016}, {11008, 4096}, {4096, 32000}, {2048, 6144},
                                                             {8192, 57344}, {28672, 8192}, {256, 1000}, {100, 36}, {33, 5}, {1024, 12}, {300, 1021}})
                check_matrix(kn.first, kn.second, bits, sms);
    }
    std::printf(g_fail ? "layout test: %d failures\n" : "layout test: ok\n", g_fail);
    return g_fail ? 1 : 0;
}
