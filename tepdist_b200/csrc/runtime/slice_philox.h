// Host-side shard slicing and sharded deterministic initialisation.
//
// Reference parity: SliceUtils (xla/pjrt/slice_utils.h: N-d strided shard copy for multi-level and
// stride-on-dim splits; GetSliceStartOffsetOnSrc returns (offset, len) runs) and the Philox initialisers
// (xla/pjrt/initializers.{h,cc}, fill_philox_random.h): every worker fills ONLY its shard, bit-identical to the
// corresponding slice of an unsharded initialisation.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "ir.h"

namespace tepdist {

// Contiguous (element offset, length) runs, in shard order, that make up the shard of `shape` selected by
// applying every split level in `levels` with the per-level index `ids` (levels[i] is glue/partial => no cut).
std::vector<std::pair<int64_t, int64_t>> SliceRuns(const std::vector<int64_t>& shape, const std::vector<DimStrategy>& levels,
                                                   const std::vector<int>& ids);
std::vector<int64_t> ShardShape(const std::vector<int64_t>& shape, const std::vector<DimStrategy>& levels);
// Copy the shard out of a full row-major buffer (elem_size bytes per element).
void SliceCopy(const uint8_t* src, uint8_t* dst, int64_t elem_size, const std::vector<int64_t>& shape,
               const std::vector<DimStrategy>& levels, const std::vector<int>& ids);

// Philox4x32-10 stream keyed by `seed`; element i of the logical tensor depends only on (seed, i).
void PhiloxFill(const std::string& kind, uint64_t seed, int64_t offset, int64_t n, float mean, float stddev, float lo, float hi,
                float* out);
// Fill exactly the shard described by (shape, levels, ids): equals SliceCopy of a full PhiloxFill.
std::vector<float> PhiloxFillShard(const std::string& kind, uint64_t seed, const std::vector<int64_t>& shape,
                                   const std::vector<DimStrategy>& levels, const std::vector<int>& ids, float mean, float stddev,
                                   float lo, float hi);

}  // namespace tepdist
