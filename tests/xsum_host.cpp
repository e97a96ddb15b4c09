// Host-side driver of rlfluidcontrol_b200/csrc/exact_sum.cuh for tests/test_exact_sum.py (no GPU needed):
// the same segment summaries and serial pass the CUDA kernels run, executed on the CPU, so the exactness of
// the re-phrased Field.sum can be checked against the plain serial loop on arbitrary data.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../rlfluidcontrol_b200/csrc/exact_sum.cuh"

using namespace rlfc::xsum;

extern "C" {

// the reference loop (Field.pde:311-318): one float accumulator, elements in order
float xs_serial(const float* a, long n) {
  volatile float s = 0.f;
  for (long k = 0; k < n; k++) s = s + a[k];
  return s;
}

// stats[0..2] = segments summarised as one table / split / serial, stats[3] = summaries rejected by the serial pass,
// stats[4] = stretches crossed with one composed table, stats[5] = stretches walked slot by slot.
// pred_noise perturbs the predicted accumulator (relative) to exercise wrong predictions.
float xs_parallel(const float* a, long n, double pred_noise, long* stats) {
  const long nseg = (n + kSeg - 1) / kSeg;
  std::vector<double> segsum(nseg, 0.0);
  for (long g = 0; g < nseg; g++) {
    double t = 0;
    for (long k = g * kSeg; k < n && k < (g + 1) * kSeg; k++) t += (double)a[k];
    segsum[g] = t;
  }
  std::vector<uint32_t> slots((size_t)nseg * kSlotWords);
  double pred = 0;
  for (long g = 0; g < nseg; g++) {
    const float* seg = a + g * kSeg;
    const int cnt = (int)((n - g * kSeg) < kSeg ? (n - g * kSeg) : kSeg);
    const double p = pred * (1.0 + pred_noise * ((g * 2654435761u) % 1000 / 500.0 - 1.0));
    build_segment([&](int k) { return seg[k]; }, cnt, p, &slots[(size_t)g * kSlotWords]);
    pred += segsum[g];
  }
  for (int k = 0; k < 6; k++) stats[k] = 0;
  auto redo = [&](uint32_t& bits, long g) {
    volatile float s = u2f(bits);
    for (long k = g * kSeg; k < n && k < (g + 1) * kSeg; k++) s = s + a[k];
    bits = f2u(s);
  };
  uint32_t bits = 0;   // s = +0.f
  // batches of 32 slots, exactly as k_xsum_tables / k_xsum_chain organise them
  for (long b0 = 0; b0 < nseg; b0 += 32) {
    uint32_t w[32][kSlotWords];
    bool plain[32];
    for (int k = 0; k < 32; k++) {
      if (b0 + k < nseg) std::memcpy(w[k], &slots[(size_t)(b0 + k) * kSlotWords], sizeof(w[k]));
      else { std::memset(w[k], 0, sizeof(w[k])); w[k][0] = kOne; w[k][1] = kAnyKey; }
      stats[w[k][0]] += (b0 + k < nseg);
      plain[k] = w[k][0] == kOne;
      if (plain[k]) {                    // stretch table in words 8..14 (sequential composition; the device scans a tree)
        normalise_table(w[k] + 1);
        std::memcpy(w[k] + 8, w[k] + 1, 7 * sizeof(uint32_t));
        if (k > 0 && plain[k - 1]) compose_tables(w[k - 1] + 8, w[k] + 8);
      }
    }
    int cur = 0;
    bool fresh = true;
    while (cur < 32) {
      int f = cur;
      while (f < 32 && plain[f]) f++;
      if (f > cur) {
        bool crossed = false;
        if (fresh) {
          const uint32_t* t = w[f - 1] + 8;
          bool ok = true;
          const uint32_t nb = apply_table(bits, t[0], (int32_t)t[1], (int32_t)t[2], (int32_t)t[3], (int32_t)t[4], (int32_t)t[5],
                                          (int32_t)t[6], ok);
          if (ok) { bits = nb; crossed = true; stats[4]++; }
        }
        if (!crossed) {
          stats[5]++;
          int g = cur;
          for (; g < f; g++) {
            const uint32_t* t = w[g] + 1;
            bool ok = true;
            const uint32_t nb = apply_table(bits, t[0], (int32_t)t[1], (int32_t)t[2], (int32_t)t[3], (int32_t)t[4], (int32_t)t[5],
                                            (int32_t)t[6], ok);
            if (!ok) break;
            bits = nb;
          }
          if (g < f) { stats[3]++; redo(bits, b0 + g); cur = g + 1; fresh = false; continue; }
        }
      }
      if (f < 32) {
        if (!apply_segment(bits, w[f])) { if (w[f][0] != kSerial) stats[3]++; redo(bits, b0 + f); }
      }
      cur = f + 1;
      fresh = true;
    }
  }
  return u2f(bits);
}
}
