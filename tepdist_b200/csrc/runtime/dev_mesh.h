// Device-mesh addressing (reference: xla/pjrt/dev_id_util.{h,cc} — SplitId, DevGroup, CommDevManager).
//
// A tensor shard / task is addressed by a multi-level split id `ids[ordinal]`.  Levels flagged share_dev (the
// micro-batch level) are time-multiplexed on the same device and contribute base 0; the other levels form a
// mixed-radix device id in `placement_layout` order (outermost first — the stage level is rotated to the front so
// pipeline stages map to contiguous device blocks).
#pragma once
#include <string>
#include <vector>

namespace tepdist {

struct SplitId {
  std::vector<int> ids;  // one per split ordinal
  int micro_id(const std::vector<bool>& share_dev) const;
  int stage_id(int stage_ordinal) const { return stage_ordinal >= 0 && stage_ordinal < (int)ids.size() ? ids[stage_ordinal] : 0; }
  std::string spmd_str() const;
};

struct DevGroup {
  int ordinal;
  std::vector<int> devices;  // global device ids, ordered by rank in group
};

class CommDevManager {
 public:
  void Build(const std::vector<int>& split_nums, const std::vector<bool>& share_dev, std::vector<int> placement_layout,
             int num_workers, int devs_per_worker);
  int total_devices() const { return total_; }
  int GlobalDevice(const SplitId& id) const;
  int WorkerOf(int global_dev) const { return devs_per_worker_ > 0 ? global_dev / devs_per_worker_ : 0; }
  int LocalDevice(int global_dev) const { return devs_per_worker_ > 0 ? global_dev % devs_per_worker_ : global_dev; }
  std::vector<int> Coords(int global_dev) const;               // per ordinal (0 for share_dev levels)
  DevGroup GroupOf(int global_dev, int ordinal) const;         // the communicator this device uses at `ordinal`
  int RankInGroup(int global_dev, int ordinal) const;
  std::vector<DevGroup> AllGroups(int ordinal) const;
  bool GroupSpansWorkers(const DevGroup& g) const;
  const std::vector<int>& split_nums() const { return split_nums_; }
  const std::vector<bool>& share_dev() const { return share_dev_; }
  std::string Describe() const;

 private:
  std::vector<int> split_nums_;
  std::vector<bool> share_dev_;
  std::vector<int> layout_;
  std::vector<int> base_;  // per ordinal; 0 for shared levels
  int total_ = 1, num_workers_ = 1, devs_per_worker_ = 1;
};

}  // namespace tepdist
