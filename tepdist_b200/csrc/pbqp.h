// Exact solver for "pick one option per node, pay node cost + pairwise edge cost" problems.
//
// This is the strategy-selection core of the SPMD planner.  Degree-0/1/2 reductions solve every tree-/chain-
// shaped part optimally by dynamic programming (what the reference calls "DP inside cones"); the irreducible
// core is solved by branch & bound (the reference's cone ILP handed to COIN-OR CBC), with a
// time limit after which the greedy RN rule finishes the job (mirroring ILP_TIME_LIMIT + fallback).
#pragma once
#include <map>
#include <vector>

namespace tepdist {

class PBQP {
 public:
  using Vec = std::vector<double>;
  using Mat = std::vector<Vec>;  // [option of u][option of v]

  int AddNode(const Vec& costs);
  void AddEdge(int u, int v, const Mat& m);  // accumulates if the edge exists
  int num_nodes() const { return (int)cost_.size(); }

  struct Result {
    std::vector<int> choice;
    double cost = 0;
    bool optimal = true;
    int core_nodes = 0;       // nodes left after reductions (size of the "ILP" part)
    int reduced_nodes = 0;    // nodes eliminated by DP reductions
    long bb_nodes = 0;
  };
  Result Solve(double time_limit_s = 30.0);
  double Evaluate(const std::vector<int>& choice) const;

 private:
  std::vector<Vec> cost_;
  std::vector<std::map<int, Mat>> adj_;  // adj_[u][v] = matrix indexed [opt_u][opt_v]
};

}  // namespace tepdist
