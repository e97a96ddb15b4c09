// Host-side driver of rlfluidcontrol_b200/csrc/exact_sum.cuh for tests/test_exact_sum.py (no GPU needed):
// the same segment summaries and serial pass the CUDA kernels run, executed on the CPU, so the exactness of
// the re-phrased Field.sum can be checked against the plain serial loop on arbitrary data.
#include <cstdint>
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

// stats[0..2] = segments summarised as one table / split / serial, stats[3] = summaries rejected by the serial pass.
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
  for (int k = 0; k < 4; k++) stats[k] = 0;
  uint32_t bits = 0;   // s = +0.f
  for (long g = 0; g < nseg; g++) {
    const uint32_t* slot = &slots[(size_t)g * kSlotWords];
    stats[slot[0]]++;
    if (!apply_segment(bits, slot)) {
      if (slot[0] != kSerial) stats[3]++;
      volatile float s = u2f(bits);
      for (long k = g * kSeg; k < n && k < (g + 1) * kSeg; k++) s = s + a[k];
      bits = f2u(s);
    }
  }
  return u2f(bits);
}
}
