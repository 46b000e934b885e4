// Host harness (test infrastructure): runs the product's row expansion (putslam_b200/csrc/lc_tc.cuh: expand_row, decode_row)
// on the CPU and evaluates the int8 dot products straight from the canonical K-major tile layout the tensor cores read
// (8 x 16 B core matrices, LBO between K chunks, SBO between 8-row groups).  Prints "ok" or the first violation.
//   nvcc -std=c++17 -o tc_encode_host tc_encode_host.cu   (no GPU needed: only host code runs)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "../putslam_b200/csrc/lc_tc.cuh"
#include "../putslam_b200/csrc/common.cuh"
using namespace pslam::tc;

static int8_t elem(const std::vector<uint8_t>& tile, int row, int k) {   // element k of operand row `row`
    return (int8_t)tile[(size_t)(row >> 3) * kSBO + (size_t)(k >> 4) * kLBO + (size_t)(row & 7) * 16 + (size_t)(k & 15)];
}
static int dot(const std::vector<uint8_t>& a, int ra, const std::vector<uint8_t>& b, int rb) {
    int s = 0;
    for (int k = 0; k < kRowBytes; ++k) s += (int)elem(a, ra, k) * (int)elem(b, rb, k);
    return s;
}
int main() {
    uint64_t st = 0x9e3779b97f4a7c15ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 16); };
    const int NQ = 256, NT = 256, nq_valid = 232, nt_valid = 200;          // the tails are padding rows
    std::vector<uint8_t> q(2 * kTileBytes), t(2 * kTileBytes);
    std::vector<uint32_t> qw(NQ * 8), tw(NT * 8);
    for (auto& x : qw) x = rnd();
    for (auto& x : tw) x = rnd();
    for (int w = 0; w < 8; ++w) { tw[5 * 8 + w] = qw[3 * 8 + w]; tw[9 * 8 + w] = qw[3 * 8 + w]; tw[77 * 8 + w] = ~qw[100 * 8 + w]; }  // distance 0 twice, 256
    for (int r = 0; r < NQ; ++r) {
        uint32_t w[8];
        for (int i = 0; i < 8; ++i) w[i] = r < nq_valid ? qw[r * 8 + i] : 0;
        expand_row(w, r < nq_valid, q.data() + (r >> 7) * kTileBytes, r & 127, 255 - r, true);
    }
    for (int r = 0; r < NT; ++r) {
        uint32_t w[8];
        for (int i = 0; i < 8; ++i) w[i] = tw[r * 8 + i];
        // the map stores rows re-encoded (ham256_encode): the producer undoes that
        ham256_encode(w);
        pslam::tc::decode_row(w);
        for (int i = 0; i < 8; ++i) if (w[i] != tw[r * 8 + i]) { printf("decode_row does not invert ham256_encode (row %d word %d)\n", r, i); return 1; }
        if (r >= nt_valid) for (int i = 0; i < 8; ++i) w[i] = 0;
        expand_row(w, r < nt_valid, t.data() + (r >> 7) * kTileBytes, r & 127, 255 - r, false);
    }
    int min_valid = 1 << 30, max_padding = -(1 << 30);
    for (int i = 0; i < NQ; ++i)
        for (int j = 0; j < NT; ++j) {
            const int acc = dot(q, i, t, j);
            if (i < nq_valid && j < nt_valid) {
                int ham = 0;
                for (int w = 0; w < 8; ++w) ham += __builtin_popcount(qw[i * 8 + w] ^ tw[j * 8 + w]);
                const int want = 512 * (128 - ham) + (255 - j) + (255 - i);
                if (acc != want) { printf("accumulator (%d, %d) = %d, expected %d\n", i, j, acc, want); return 1; }
                if ((acc >> kStepShift) != 128 - ham || (acc & 511) != (255 - j) + (255 - i)) { printf("field decode (%d, %d)\n", i, j); return 1; }
                if (acc < min_valid) min_valid = acc;
            } else if (acc > max_padding) max_padding = acc;
        }
    if (max_padding >= min_valid || max_padding >= 512 * (128 - 256)) { printf("a padding row can win: %d vs %d\n", max_padding, min_valid); return 1; }
    // ordering: the row maximum is the lowest distance, then the lowest t (rows 5 and 9 both hold query 3: 5 must win)
    int best = -(1 << 30), bj = -1;
    for (int j = 0; j < nt_valid; ++j) { const int a = dot(q, 3, t, j); if (a > best) { best = a; bj = j; } }
    if (bj != 5) { printf("tie-break along a row: %d\n", bj); return 1; }
    printf("ok %d %d\n", min_valid, max_padding);
    return 0;
}
