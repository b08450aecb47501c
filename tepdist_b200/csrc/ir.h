// Planner IR (C++ side): the same schema as tepdist_b200/ir.py, plus the sharding annotations the planner
// writes onto it.
//
// Reference parity:
//   DimDistSpec / DistSpec      xla/service/parallel/dist_spec.h:36-227
//   DimStrategy / HLOStrategy   xla/service/parallel/hlo_strategy_spec.{h,cc}
//   OpMetadata op_group/backward, module split metadata (split_nums, share_dev_flags, placement_layout)
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <utility>
#include <variant>
#include <vector>

namespace tepdist {

using Attr = std::variant<int64_t, double, std::string, std::vector<int64_t>, bool>;

struct TensorType {
  std::vector<int64_t> dims;
  std::string dtype;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : dims) n *= d;
    return n;
  }
  int64_t elem_bytes() const {
    if (dtype == "bf16" || dtype == "f16") return 2;
    if (dtype == "f32" || dtype == "i32") return 4;
    if (dtype == "i64") return 8;
    return 1;
  }
  int64_t bytes() const { return numel() * elem_bytes(); }
  int rank() const { return (int)dims.size(); }
};

struct ValueRef {
  int node = -1;
  int idx = 0;
  bool operator<(const ValueRef& o) const { return node != o.node ? node < o.node : idx < o.idx; }
  bool operator==(const ValueRef& o) const { return node == o.node && idx == o.idx; }
};

// How one tensor is laid out across the `num` devices of ONE mesh level (split ordinal).
//   glue      : dim < 0 && !partial  -> replicated / undecided (the reference's "Glue")
//   split     : dim >= 0             -> dimension `dim` is cut in `num` pieces.  `stride` (elements along
//               `dim`, 0 = whole extent) makes the split layout-aware: within every block of `stride`
//               elements shard k owns the k-th 1/num  (reference: stride_on_dim, for reshape-merged dims,
//               e.g. the fused qkv projection whose last dim is [3][H][D]).
//   partial   : every device holds a full-shape addend; the true value is the reduction over devices.
struct DimStrategy {
  int dim = -1;
  int num = 1;
  int64_t stride = 0;
  bool partial = false;
  int reduce_kind = 0;  // 0 sum, 1 max, 2 min, 3 prod (partial only)

  static DimStrategy Glue() { return DimStrategy(); }
  static DimStrategy Split(int d, int n, int64_t stride = 0) {
    DimStrategy s;
    s.dim = d;
    s.num = n;
    s.stride = stride;
    return s;
  }
  static DimStrategy Partial(int n, int kind = 0) {
    DimStrategy s;
    s.num = n;
    s.partial = true;
    s.reduce_kind = kind;
    return s;
  }
  bool is_glue() const { return dim < 0 && !partial; }
  bool is_split() const { return dim >= 0; }
  bool operator==(const DimStrategy& o) const {
    return dim == o.dim && stride == o.stride && partial == o.partial && (is_glue() || num == o.num);
  }
  bool operator!=(const DimStrategy& o) const { return !(*this == o); }
  bool operator<(const DimStrategy& o) const {
    if (partial != o.partial) return partial < o.partial;
    if (dim != o.dim) return dim < o.dim;
    return stride < o.stride;
  }
  std::string str() const;
  // Normalised stride for a concrete shape (0 -> full extent); false if the split does not divide evenly.
  bool Valid(const TensorType& t) const;
  int64_t EffStride(const TensorType& t) const { return stride > 0 ? stride : t.dims[dim]; }
  // Row-major element stride of one period of the split (reference: stride_on_elements).
  int64_t StrideOnElements(const TensorType& t) const;
  // Re-derive (dim, stride) on a reshaped tensor; returns Glue if not expressible (reference ApplyToShape).
  DimStrategy ApplyToShape(const TensorType& from, const TensorType& to) const;
};

// Per value: one entry per split ordinal (mesh level) + pipeline stage.
struct DistSpec {
  std::vector<DimStrategy> levels;
  int stage = -1;
  std::string str() const;
};

struct Node {
  int id = 0;
  std::string op;
  std::vector<ValueRef> inputs;
  std::vector<TensorType> outputs;
  std::map<std::string, Attr> attrs;
  std::string name;
  int group = -1;
  bool backward = false;
  int stage = -1;
  std::vector<DistSpec> dist;  // per output (planner result)

  int64_t attr_i(const std::string& k, int64_t def = 0) const;
  double attr_f(const std::string& k, double def = 0) const;
  bool attr_b(const std::string& k, bool def = false) const;
  std::string attr_s(const std::string& k, const std::string& def = "") const;
  std::vector<int64_t> attr_v(const std::string& k) const;
  bool has(const std::string& k) const { return attrs.count(k) > 0; }
};

struct Use {
  int node;
  int operand;
};

class Graph {
 public:
  std::string name;
  std::vector<Node> nodes;
  std::vector<ValueRef> outputs;
  std::map<int, ValueRef> updates;  // variable node id -> updated value (input/output alias)
  // module-level split metadata (reference hlo_module.h diff: split_nums_, share_dev_flags_, placement_layout_)
  std::vector<int> split_nums;
  std::vector<bool> share_dev;
  std::vector<int> placement_layout;
  int stage_split_ordinal = -1;
  std::map<std::string, std::string> meta;

  int AddNode(const std::string& op, const std::vector<ValueRef>& inputs, const std::vector<TensorType>& outs,
              const std::map<std::string, Attr>& attrs, const std::string& name, int group, bool backward);
  const TensorType& type(ValueRef v) const { return nodes[v.node].outputs[v.idx]; }
  // value -> uses (rebuilt on demand)
  const std::vector<Use>& users(ValueRef v) const;
  void InvalidateUsers() { users_valid_ = false; }
  int record_split(int num, bool share) {
    split_nums.push_back(num);
    share_dev.push_back(share);
    return (int)split_nums.size() - 1;
  }
  std::string Dump(bool with_dist = true) const;

 private:
  mutable bool users_valid_ = false;
  mutable std::map<ValueRef, std::vector<Use>> users_;
  mutable std::vector<Use> empty_;
};

bool IsSource(const std::string& op);
bool IsVariable(const std::string& op);           // parameter / state
bool IsComputeIntensive(const std::string& op);   // dot / conv family (has split proposals, never Glue)
bool IsCollective(const std::string& op);
// FLOPs of one node at full (unsharded) shape (reference: PerfUtils, performance_utils.cc:37-136).
double NodeFlops(const Graph& g, const Node& n);

}  // namespace tepdist
