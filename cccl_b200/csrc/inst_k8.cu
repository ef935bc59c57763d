// onesweep instantiations for 8-byte keys (u64 / i64 / f64).  Index 0 of each table is the default configuration; the others are kept
// for A/B measurement (tools/sweep.py) and are all covered by the parity tests.
#include "inst.cuh"

namespace b200rs
{
using K = uint64_t;
#define C(VB, NT, IPT, MINB) make_config<K, VB, NT, IPT, RANK_BALLOT, MINB>()
#define O(VB, NT, IPT, MINB, OPT) make_config<K, VB, NT, IPT, RANK_BALLOT, MINB, OPT>()
#define OB(VB, NT, IPT, MINB, OPT) make_config_with_bucket<K, VB, NT, IPT, RANK_BALLOT, MINB, OPT>()
#define T(VB, NT, IPT, MINB, LBW) make_tma_config<K, VB, NT, IPT, MINB, LBW>()

static const OnesweepConfig cfg_v0[] = {
  OB(0, 256, 20, 4, 7),
  O(0, 256, 24, 3, 7),
  C(0, 256, 24, 3),
  O(0, 256, 32, 2, 7),
  T(0, 256, 24, 3, 4)
};
static const OnesweepConfig cfg_v1[] = {
  O(1, 256, 24, 3, 7),
  C(1, 256, 24, 3)
};
static const OnesweepConfig cfg_v2[] = {
  O(2, 256, 24, 3, 7),
  C(2, 256, 24, 3)
};
static const OnesweepConfig cfg_v4[] = {
  OB(4, 256, 20, 3, 7),
  C(4, 256, 24, 3),
  O(4, 256, 24, 3, 7),
  T(4, 256, 14, 3, 4)
};
static const OnesweepConfig cfg_v8[] = {
  OB(8, 256, 16, 3, 7),
  C(8, 256, 16, 3)
};
static const OnesweepConfig cfg_v16[] = {
  O(16, 256, 10, 3, 7),
  C(16, 256, 10, 3)
};

#define B200RS_TABLE(arr)                     \
  *count = int(sizeof(arr) / sizeof(arr[0])); \
  return arr

const OnesweepConfig* onesweep_configs_k8(int value_bytes, int* count)
{
  switch (value_bytes)
  {
    case 0: B200RS_TABLE(cfg_v0);
    case 1: B200RS_TABLE(cfg_v1);
    case 2: B200RS_TABLE(cfg_v2);
    case 4: B200RS_TABLE(cfg_v4);
    case 8: B200RS_TABLE(cfg_v8);
    case 16: B200RS_TABLE(cfg_v16);
    default: *count = 0; return nullptr;
  }
}
} // namespace b200rs
