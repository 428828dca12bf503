// parse_emul.cpp -- TEST INFRASTRUCTURE.  Runs the stage functions of
// gt-scaffold_b200/csrc/gtsb_parse_core.h (the bodies of the CUDA kernels in
// gtsb_parse.cu) as plain loops on the host, in the order the device driver
// launches them, so that the token rules can be compared with the compiled
// reference where there is no GPU.  Nothing in the product links this file.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../gt-scaffold_b200/csrc/gtsb_parse_core.h"

using namespace gtsbp;

namespace {
struct State {
  std::vector<uint32_t> root, ctg, num_pairs;
  std::vector<int32_t> dist;
  std::vector<float> std_dev;
  std::vector<uint8_t> flags;
} g;
}  // namespace

extern "C" {

// returns 0; *irregular != 0: nothing parsed.  `order` 0: threads in index order, 1: reversed
// (the result must not depend on which thread runs first)
int emul_parse_de(uint64_t V, const char *names, const uint64_t *name_off, const char *text_in, uint64_t n,
                  int order, uint64_t *nof_records, uint32_t *irregular) {
  uint32_t irr = 0;
  *nof_records = 0;
  uint64_t cap = 2;
  while (cap < 2 * V) cap <<= 1;
  std::vector<uint64_t> slots(cap, 0);
  const NameTable t{names, name_off, slots.data(), cap - 1};
  for (uint64_t k = 0; k < V; k++) table_insert(t, (uint32_t) (order ? V - 1 - k : k), &irr);
  if (irr) {
    *irregular = irr;
    return 0;
  }
  // 8-byte aligned copy, as the device buffer is
  std::vector<uint64_t> aligned(n / 8 + 2, 0);
  char *text = (char *) aligned.data();
  memcpy(text, text_in, n);

  const uint64_t nchunks = (n + CHUNK - 1) / CHUNK;
  std::vector<uint8_t> cnt(nchunks + 1, 0);
  for (uint64_t k = 0; k < nchunks; k++) {
    const uint64_t i = order ? nchunks - 1 - k : k;
    cnt[i] = (uint8_t) chunk_newlines(text, n, i, &irr);
  }
  std::vector<uint32_t> first(nchunks + 1, 0);
  for (uint64_t i = 0; i < nchunks; i++) first[i + 1] = first[i] + cnt[i];
  const uint32_t newlines = first[nchunks];
  const bool open_end = n != 0 && text_in[n - 1] != '\n';
  const uint64_t nlines = (uint64_t) newlines + (open_end ? 1 : 0);
  std::vector<uint64_t> line_end(nlines + 1, 0);
  for (uint64_t k = 0; k < nchunks; k++) {
    const uint64_t i = order ? nchunks - 1 - k : k;
    if (cnt[i]) chunk_line_ends(text, n, i, first[i], line_end.data());
  }
  if (open_end) line_end[newlines] = n;

  Records out{};
  std::vector<uint32_t> line_cnt(nlines + 1, 0), line_off(nlines + 2, 0);
  for (uint64_t k = 0; k < nlines; k++) {
    const uint64_t l = order ? nlines - 1 - k : k;
    line_cnt[l] = walk_line<false>(text, l ? line_end[l - 1] : 0, line_end[l], t, out, 0, &irr);
  }
  for (uint64_t l = 0; l < nlines; l++) line_off[l + 1] = line_off[l] + line_cnt[l];
  if (irr) {
    *irregular = irr;
    return 0;
  }
  const uint64_t R = line_off[nlines];
  g.root.assign(R, 0xDEADBEEF);
  g.ctg.assign(R, 0xDEADBEEF);
  g.num_pairs.assign(R, 0xDEADBEEF);
  g.dist.assign(R, 0);
  g.std_dev.assign(R, 0);
  g.flags.assign(R, 0xFF);
  out = Records{g.root.data(), g.ctg.data(), g.dist.data(), g.std_dev.data(), g.num_pairs.data(), g.flags.data()};
  for (uint64_t k = 0; k < nlines; k++) {
    const uint64_t l = order ? nlines - 1 - k : k;
    if (line_off[l + 1] != line_off[l])
      walk_line<true>(text, l ? line_end[l - 1] : 0, line_end[l], t, out, line_off[l], &irr);
  }
  *nof_records = R;
  *irregular = irr;
  return 0;
}

void emul_fetch(uint32_t *root, uint32_t *ctg, int32_t *dist, float *std_dev, uint8_t *flags, uint32_t *num_pairs) {
  const size_t R = g.root.size();
  if (R == 0) return;
  memcpy(root, g.root.data(), R * 4);
  memcpy(ctg, g.ctg.data(), R * 4);
  memcpy(dist, g.dist.data(), R * 4);
  memcpy(std_dev, g.std_dev.data(), R * 4);
  memcpy(flags, g.flags.data(), R);
  memcpy(num_pairs, g.num_pairs.data(), R * 4);
}

// `.astat` text: astat / copy_num (V values each) are updated in place
int emul_parse_astat(uint64_t V, const char *names, const uint64_t *name_off, const char *text_in, uint64_t n,
                     int order, float *astat, float *copy_num, uint32_t *irregular) {
  uint32_t irr = 0;
  uint64_t cap = 2;
  while (cap < 2 * V) cap <<= 1;
  std::vector<uint64_t> slots(cap, 0);
  const NameTable t{names, name_off, slots.data(), cap - 1};
  for (uint64_t k = 0; k < V; k++) table_insert(t, (uint32_t) (order ? V - 1 - k : k), &irr);
  if (irr) {
    *irregular = irr;
    return 0;
  }
  std::vector<uint64_t> aligned(n / 8 + 2, 0);
  char *text = (char *) aligned.data();
  memcpy(text, text_in, n);
  const uint64_t nchunks = (n + CHUNK - 1) / CHUNK;
  std::vector<uint8_t> cnt(nchunks + 1, 0);
  for (uint64_t i = 0; i < nchunks; i++) cnt[i] = (uint8_t) chunk_newlines(text, n, i, &irr);
  std::vector<uint32_t> first(nchunks + 1, 0);
  for (uint64_t i = 0; i < nchunks; i++) first[i + 1] = first[i] + cnt[i];
  const uint32_t newlines = first[nchunks];
  const bool open_end = n != 0 && text_in[n - 1] != '\n';
  const uint64_t nlines = (uint64_t) newlines + (open_end ? 1 : 0);
  std::vector<uint64_t> line_end(nlines + 1, 0);
  for (uint64_t i = 0; i < nchunks; i++)
    if (cnt[i]) chunk_line_ends(text, n, i, first[i], line_end.data());
  if (open_end) line_end[newlines] = n;

  std::vector<uint64_t> last(V + 1, 0);
  std::vector<float> a(astat, astat + V), cn(copy_num, copy_num + V);
  for (uint64_t k = 0; k < nlines; k++) {
    const uint64_t l = order ? nlines - 1 - k : k;
    walk_astat_line<false>(text, l ? line_end[l - 1] : 0, line_end[l], t, last.data(), a.data(), cn.data(), &irr);
  }
  *irregular = irr;
  if (irr) return 0;
  for (uint64_t k = 0; k < nlines; k++) {
    const uint64_t l = order ? nlines - 1 - k : k;
    walk_astat_line<true>(text, l ? line_end[l - 1] : 0, line_end[l], t, last.data(), a.data(), cn.data(), &irr);
  }
  if (V) {
    memcpy(astat, a.data(), V * 4);
    memcpy(copy_num, cn.data(), V * 4);
  }
  return 0;
}

// the float rule alone: 0 ok (*out set), else an IRR_* bit
uint32_t emul_canonical_float(const char *s, uint32_t n, float *out) { return canonical_float(s, 0, n, out); }

}  // extern "C"
