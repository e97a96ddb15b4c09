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

// stats[0..2] = segments summarised as one table / split / serial, stats[3] = batch records rejected by the serial pass,
// stats[4] = batches crossed by their record, stats[5] = batches redone as genuine additions.
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
  // batches of 32 summaries, exactly as k_xsum_tables / k_xsum_chain organise them: a record of entries
  // "table, then one float addition" per batch; any entry that does not apply -> the batch is walked summary by summary
  const uint32_t ident[7] = {kAnyKey, 0u, 0u, (uint32_t)INT32_MIN, (uint32_t)INT32_MAX, (uint32_t)INT32_MIN, (uint32_t)INT32_MAX};
  for (long b0 = 0; b0 < nseg; b0 += 32) {
    uint32_t w[32][kSlotWords];
    for (int k = 0; k < 32; k++) {
      if (b0 + k < nseg) std::memcpy(w[k], &slots[(size_t)(b0 + k) * kSlotWords], sizeof(w[k]));
      else { std::memset(w[k], 0, sizeof(w[k])); w[k][0] = kOne; w[k][1] = kAnyKey; }
      stats[w[k][0]] += (b0 + k < nseg);
      normalise_table(w[k] + 1);
      if (w[k][0] == kSplit) normalise_table(w[k] + kSlotB);
    }
    // record (sequential composition; the device scans a tree); a run of consecutive serial summaries is one entry
    struct Entry { uint32_t t[7]; float raw[3]; int serial_slot, serial_run; };
    std::vector<Entry> rec;
    uint32_t C[7];
    std::memcpy(C, ident, sizeof(C));
    for (int k = 0; k < 32; k++) {
      if (w[k][0] == kOne) {
        uint32_t g[7]; std::memcpy(g, w[k] + 1, sizeof(g));
        compose_tables(C, g); std::memcpy(C, g, sizeof(C));
      } else if (w[k][0] == kSplit) {
        Entry en; en.serial_slot = -1; en.serial_run = 0;
        uint32_t g[7]; std::memcpy(g, w[k] + 1, sizeof(g));
        compose_tables(C, g); std::memcpy(en.t, g, sizeof(g));
        for (int r = 0; r < 3; r++) en.raw[r] = u2f(w[k][kSlotRaw + r]);
        std::memcpy(C, w[k] + kSlotB, sizeof(C));
        rec.push_back(en);
      } else if (k > 0 && w[k - 1][0] == kSerial) {
        rec.back().serial_run++;
      } else {
        Entry en; en.raw[0] = en.raw[1] = en.raw[2] = -0.f;
        std::memcpy(en.t, C, sizeof(C));
        en.serial_slot = k; en.serial_run = 1;
        std::memcpy(C, ident, sizeof(C));
        rec.push_back(en);
      }
    }
    { Entry en; std::memcpy(en.t, C, sizeof(C)); en.raw[0] = en.raw[1] = en.raw[2] = -0.f; en.serial_slot = -1; en.serial_run = 0; rec.push_back(en); }
    bool redo_batch = rec.size() > 15;
    if (!redo_batch) {
      const uint32_t start = bits;
      bool ok = true;
      for (const Entry& en : rec) {
        bits = apply_table(bits, en.t[0], (int32_t)en.t[1], (int32_t)en.t[2], (int32_t)en.t[3], (int32_t)en.t[4], (int32_t)en.t[5],
                           (int32_t)en.t[6], ok);
        volatile float sv = u2f(bits); sv = sv + en.raw[0]; sv = sv + en.raw[1]; sv = sv + en.raw[2]; bits = f2u(sv);
        for (int j = 0; j < en.serial_run; j++) redo(bits, b0 + en.serial_slot + j);
      }
      if (ok) stats[4]++;
      else { bits = start; redo_batch = true; stats[3]++; }
    }
    if (redo_batch) {                    // the whole batch as genuine additions
      stats[5]++;
      for (int k = 0; k < 32; k++) redo(bits, b0 + k);
    }
  }
  return u2f(bits);
}

// The device condenses a batch with a Kogge-Stone scan of table compositions (k_xsum_tables); the serial model
// above composes left to right.  Composition must be associative INCLUDING the normalisation of dead halves, or the
// two would disagree.  Returns the number of batches whose group tables (C_k of every lane, E_k of every head)
// differ between the two orders.
long xs_tree_vs_sequential(const float* a, long n) {
  const long nseg = (n + kSeg - 1) / kSeg;
  std::vector<uint32_t> slots((size_t)nseg * kSlotWords);
  double pred = 0;
  for (long g = 0; g < nseg; g++) {
    const float* seg = a + g * kSeg;
    const int cnt = (int)((n - g * kSeg) < kSeg ? (n - g * kSeg) : kSeg);
    build_segment([&](int k) { return seg[k]; }, cnt, pred, &slots[(size_t)g * kSlotWords]);
    for (int k = 0; k < cnt; k++) pred += (double)seg[k];
  }
  const uint32_t ident[7] = {kAnyKey, 0u, 0u, (uint32_t)INT32_MIN, (uint32_t)INT32_MAX, (uint32_t)INT32_MIN, (uint32_t)INT32_MAX};
  long differ = 0;
  for (long b0 = 0; b0 < nseg; b0 += 32) {
    uint32_t X[32][7], v[32][7], type[32];
    for (int k = 0; k < 32; k++) {
      uint32_t w[kSlotWords];
      if (b0 + k < nseg) std::memcpy(w, &slots[(size_t)(b0 + k) * kSlotWords], sizeof(w));
      else { std::memset(w, 0, sizeof(w)); w[0] = kOne; w[1] = kAnyKey; }
      type[k] = w[0];
      normalise_table(w + 1);
      if (w[0] == kSplit) normalise_table(w + kSlotB);
      for (int q = 0; q < 7; q++) {
        X[k][q] = (w[0] == kSerial) ? ident[q] : w[1 + q];
        v[k][q] = (w[0] == kOne) ? w[1 + q] : (w[0] == kSplit ? w[kSlotB + q] : ident[q]);
      }
    }
    // sequential: C_k = head ? v_k : C_{k-1} o v_k ;  E_k = C_{k-1} o X_k
    uint32_t Cs[32][7], Es[32][7];
    for (int k = 0; k < 32; k++) {
      uint32_t prev[7];
      std::memcpy(prev, k ? Cs[k - 1] : ident, sizeof(prev));
      std::memcpy(Cs[k], v[k], sizeof(prev));
      if (type[k] == kOne) compose_tables(prev, Cs[k]);
      std::memcpy(Es[k], X[k], sizeof(prev));
      compose_tables(prev, Es[k]);
    }
    // tree, exactly as the kernel: dist = lane - (last head at or below the lane, else 0)
    uint32_t acc[32][7];
    std::memcpy(acc, v, sizeof(acc));
    int dist[32];
    for (int k = 0; k < 32; k++) {
      int s0 = 0;
      for (int j = 0; j <= k; j++) if (type[j] != kOne) s0 = j;
      dist[k] = k - s0;
    }
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t nxt[32][7];
      std::memcpy(nxt, acc, sizeof(acc));
      for (int k = 0; k < 32; k++)
        if (k >= o && dist[k] >= o) compose_tables(acc[k - o], nxt[k]);
      std::memcpy(acc, nxt, sizeof(acc));
    }
    bool same = true;
    for (int k = 0; k < 32; k++) {
      uint32_t E[7];
      std::memcpy(E, X[k], sizeof(E));
      compose_tables(k ? acc[k - 1] : ident, E);
      same = same && std::memcmp(acc[k], Cs[k], sizeof(E)) == 0;
      if (type[k] != kOne) same = same && std::memcmp(E, Es[k], sizeof(E)) == 0;
    }
    differ += !same;
  }
  return differ;
}
}
