// onesweep instantiations for 1-byte keys
#include "inst.cuh"

namespace b200rs
{
using K = uint8_t;

static const OnesweepConfig cfg_v0[]  = {make_config<K, 0, 512, 16, RANK_MATCH>()};
static const OnesweepConfig cfg_v1[]  = {make_config<K, 1, 512, 16, RANK_MATCH>()};
static const OnesweepConfig cfg_v2[]  = {make_config<K, 2, 512, 16, RANK_MATCH>()};
static const OnesweepConfig cfg_v4[]  = {make_config<K, 4, 512, 16, RANK_MATCH>()};
static const OnesweepConfig cfg_v8[]  = {make_config<K, 8, 512, 12, RANK_MATCH>()};
static const OnesweepConfig cfg_v16[] = {make_config<K, 16, 512, 8, RANK_MATCH>()};

#define B200RS_TABLE(arr)                            \
  *count = int(sizeof(arr) / sizeof(arr[0]));        \
  return arr

const OnesweepConfig* onesweep_configs_k1(int value_bytes, int* count)
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
