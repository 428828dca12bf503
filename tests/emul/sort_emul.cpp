// sort_emul.cpp -- TEST INFRASTRUCTURE.  The blocked bitonic sort of the general build's hub
// buckets (gt-scaffold_b200/csrc/gtsb_sort_core.h, called by k_resolve_large2 in gtsb_build.cu)
// compiled for the host: a "block" is a loop over the thread index, forwards or backwards.
#include <stdint.h>

#include <vector>

#include "../../gt-scaffold_b200/csrc/gtsb_sort_core.h"

namespace gtsbs { int emul_reverse = 0; }

extern "C" {

// ent: P x {x, y, z, w}; tag: P words or null.  Sorted in place.
void emul_blocked_bitonic(uint32_t *ent, uint32_t *tag, uint32_t P, int mode, int reverse) {
  gtsbs::emul_reverse = reverse;
  std::vector<gtsbs::Ent> s_a(gtsbs::SORT_CHUNK);
  std::vector<uint32_t> s_t(gtsbs::SORT_CHUNK);
  gtsbs::blocked_bitonic(reinterpret_cast<gtsbs::Ent *>(ent), tag, P, mode, s_a.data(), s_t.data());
}

// the network as a warp runs it over a bucket held in shared memory (k_resolve_mid)
void emul_plain_bitonic(uint32_t *ent, uint32_t *tag, uint32_t P, int mode, int reverse) {
  gtsbs::emul_reverse = reverse;
  gtsbs::plain_bitonic<gtsbs::WarpGroup>(reinterpret_cast<gtsbs::Ent *>(ent), tag, P, mode);
}

uint32_t emul_sort_chunk(void) { return gtsbs::SORT_CHUNK; }

}
