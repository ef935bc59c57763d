// onesweep instantiations for 8-byte keys with 1-, 2-, 8- and 16-byte values.  Index 0 of each table is the default configuration; the others are kept
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

static const OnesweepConfig cfg_v1[] = {
  O(1, 256, 24, 3, BASE),
  C(1, 256, 24, 3)
};
static const OnesweepConfig cfg_v2[] = {
  O(2, 256, 24, 3, BASE),
  C(2, 256, 24, 3)
};
static const OnesweepConfig cfg_v8[] = {
  OB(8, 256, 16, 3, BASE),
  O(8, 256, 16, 3, BASE | OPT_SHORT_WARP)
};
static const OnesweepConfig cfg_v16[] = {
  O(16, 256, 10, 3, BASE),
  C(16, 256, 10, 3)
};
#define B200RS_TABLE(arr)                     \
  *count = int(sizeof(arr) / sizeof(arr[0])); \
  return arr

const OnesweepConfig* onesweep_configs_k8_vx(int value_bytes, int* count)
{
  switch (value_bytes)
  {
    case 1: B200RS_TABLE(cfg_v1);
    case 2: B200RS_TABLE(cfg_v2);
    case 8: B200RS_TABLE(cfg_v8);
    case 16: B200RS_TABLE(cfg_v16);
    default: *count = 0; return nullptr;
  }
}
} // namespace b200rs
