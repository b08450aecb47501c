#include "runtime/dev_mesh.h"

#include <algorithm>
#include <set>
#include <sstream>

namespace tepdist {

int SplitId::micro_id(const std::vector<bool>& share_dev) const {
  for (size_t i = 0; i < ids.size() && i < share_dev.size(); ++i)
    if (share_dev[i]) return ids[i];
  return 0;
}
std::string SplitId::spmd_str() const {
  std::ostringstream o;
  for (size_t i = 0; i < ids.size(); ++i) o << (i ? "." : "") << ids[i];
  return o.str();
}

void CommDevManager::Build(const std::vector<int>& split_nums, const std::vector<bool>& share_dev,
                           std::vector<int> placement_layout, int num_workers, int devs_per_worker) {
  split_nums_ = split_nums;
  share_dev_ = share_dev;
  if (placement_layout.empty())
    for (int i = 0; i < (int)split_nums.size(); ++i) placement_layout.push_back(i);
  layout_ = placement_layout;
  base_.assign(split_nums.size(), 0);
  int b = 1;
  for (auto it = layout_.rbegin(); it != layout_.rend(); ++it) {  // innermost level varies fastest
    const int l = *it;
    if (share_dev_[l]) continue;
    base_[l] = b;
    b *= split_nums_[l];
  }
  total_ = b;
  num_workers_ = std::max(1, num_workers);
  devs_per_worker_ = devs_per_worker > 0 ? devs_per_worker : std::max(1, total_ / num_workers_);
}

int CommDevManager::GlobalDevice(const SplitId& id) const {
  int d = 0;
  for (size_t l = 0; l < base_.size() && l < id.ids.size(); ++l) d += base_[l] * id.ids[l];
  return d;
}
std::vector<int> CommDevManager::Coords(int dev) const {
  std::vector<int> c(base_.size(), 0);
  for (size_t l = 0; l < base_.size(); ++l)
    if (base_[l] > 0) c[l] = (dev / base_[l]) % split_nums_[l];
  return c;
}
DevGroup CommDevManager::GroupOf(int dev, int ordinal) const {
  DevGroup g;
  g.ordinal = ordinal;
  if (base_[ordinal] == 0) {
    g.devices = {dev};
    return g;
  }
  const int c = (dev / base_[ordinal]) % split_nums_[ordinal];
  const int origin = dev - c * base_[ordinal];
  for (int i = 0; i < split_nums_[ordinal]; ++i) g.devices.push_back(origin + i * base_[ordinal]);
  return g;
}
int CommDevManager::RankInGroup(int dev, int ordinal) const {
  return base_[ordinal] == 0 ? 0 : (dev / base_[ordinal]) % split_nums_[ordinal];
}
std::vector<DevGroup> CommDevManager::AllGroups(int ordinal) const {
  std::vector<DevGroup> out;
  std::set<std::vector<int>> seen;
  for (int d = 0; d < total_; ++d) {
    DevGroup g = GroupOf(d, ordinal);
    if (seen.insert(g.devices).second) out.push_back(g);
  }
  return out;
}
bool CommDevManager::GroupSpansWorkers(const DevGroup& g) const {
  for (int d : g.devices)
    if (WorkerOf(d) != WorkerOf(g.devices[0])) return true;
  return false;
}
std::string CommDevManager::Describe() const {
  std::ostringstream o;
  o << "mesh levels=[";
  for (size_t l = 0; l < split_nums_.size(); ++l)
    o << (l ? "," : "") << split_nums_[l] << (share_dev_[l] ? "(shared)" : "") << ":base" << base_[l];
  o << "] devices=" << total_ << " workers=" << num_workers_ << "x" << devs_per_worker_;
  return o.str();
}

}  // namespace tepdist
