// onesweep instantiations for 4-byte keys with 4-byte values.  Index 0 of each table is the default configuration; the others are kept
// for A/B measurement (tools/sweep.py) and are all covered by the parity tests.
#include "inst.cuh"

namespace b200rs
{
using K = uint32_t;
#define C(VB, NT, IPT, MINB) make_config<K, VB, NT, IPT, RANK_BALLOT, MINB>()
#define O(VB, NT, IPT, MINB, OPT) make_config<K, VB, NT, IPT, RANK_BALLOT, MINB, OPT>()
#define OB(VB, NT, IPT, MINB, OPT) make_config_with_bucket<K, VB, NT, IPT, RANK_BALLOT, MINB, OPT>()
#define T(VB, NT, IPT, MINB, LBW) make_tma_config<K, VB, NT, IPT, MINB, LBW>()
// 7 = FMA-pipe complement + look-back window + 16-bit counters; FAST adds the single-digit-warp short circuit and, for
// 4-byte keys, the folded table addressing (onesweep.cuh OnesweepOpt)
constexpr int BASE = OPT_FMA_NOT | OPT_LB_WINDOW | OPT_CTR16;
constexpr int FAST = BASE | OPT_SHORT_WARP | OPT_FOLD | OPT_FOLD_PTR;

static const OnesweepConfig cfg_v4[] = {
  OB(4, 256, 32, 3, FAST),
  O(4, 256, 36, 3, BASE), // the round-1 default
  C(4, 256, 36, 3),
  T(4, 256, 24, 3, 4),
  O(4, 256, 36, 3, FAST)
};
#define B200RS_TABLE(arr)                     \
  *count = int(sizeof(arr) / sizeof(arr[0])); \
  return arr

const OnesweepConfig* onesweep_configs_k4_v4(int* count)
{
  B200RS_TABLE(cfg_v4);
}
} // namespace b200rs
