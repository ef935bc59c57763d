// onesweep instantiations for 8-byte keys (u64 / i64 / f64), keys only.  Index 0 of each table is the default configuration; the others are kept
// for A/B measurement (tools/sweep.py) and are all covered by the parity tests.
#include "inst.cuh"

namespace b200rs
{
using K = uint64_t;
#define C(VB, NT, IPT, MINB) make_config<K, VB, NT, IPT, RANK_BALLOT, MINB>()
#define O(VB, NT, IPT, MINB, OPT) make_config<K, VB, NT, IPT, RANK_BALLOT, MINB, OPT>()
#define OB(VB, NT, IPT, MINB, OPT) make_config_with_bucket<K, VB, NT, IPT, RANK_BALLOT, MINB, OPT>()
#define T(VB, NT, IPT, MINB, LBW) make_tma_config<K, VB, NT, IPT, MINB, LBW>()
// 7 = FMA-pipe complement + look-back window + 16-bit counters; FAST adds the single-digit-warp short circuit and, for
// 4-byte keys, the folded table addressing (onesweep.cuh OnesweepOpt)
constexpr int BASE = OPT_FMA_NOT | OPT_LB_WINDOW | OPT_CTR16;
constexpr int FAST = BASE | OPT_SHORT_WARP | OPT_FOLD | OPT_FOLD_PTR;

static const OnesweepConfig cfg_v0[] = {
  OB(0, 256, 20, 4, BASE),
  O(0, 256, 24, 3, BASE),
  C(0, 256, 24, 3),
  T(0, 256, 24, 3, 4),
  O(0, 256, 20, 4, BASE | OPT_SHORT_WARP) // single-digit warps skip ranking: -2 % on uniform keys, kept for A/B
};
#define B200RS_TABLE(arr)                     \
  *count = int(sizeof(arr) / sizeof(arr[0])); \
  return arr

const OnesweepConfig* onesweep_configs_k8_v0(int* count)
{
  B200RS_TABLE(cfg_v0);
}
} // namespace b200rs
