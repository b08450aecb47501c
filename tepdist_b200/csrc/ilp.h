// General mixed-integer linear programming: dense two-phase simplex + depth-first branch & bound.
//
// Reference parity: the reference links COIN-OR CBC (CbcModel over OsiClpSolverInterface) and uses
// it for the cone ILP and the pipeline-stage ILP with ILP_TIME_LIMIT / ILP_NUM_THREADS.  CBC is not available in
// this image, so the planner carries its own exact solver; tests cross-check it against SciPy's HiGHS.
//   minimise c^T x   s.t.  row_lo <= A x <= row_hi,   lo <= x <= hi,   x_i integer for i in integer set
#pragma once
#include <limits>
#include <string>
#include <vector>

namespace tepdist {

struct IlpModel {
  static constexpr double kInf = 1e30;
  int num_vars = 0;
  std::vector<double> obj, lo, hi;
  std::vector<char> is_int;
  struct Row {
    std::vector<int> idx;
    std::vector<double> val;
    double lo, hi;
  };
  std::vector<Row> rows;

  int AddVar(double lo_, double hi_, double cost, bool integer) {
    obj.push_back(cost); lo.push_back(lo_); hi.push_back(hi_); is_int.push_back(integer ? 1 : 0);
    return num_vars++;
  }
  void AddRow(const std::vector<int>& idx, const std::vector<double>& val, double lo_, double hi_) {
    rows.push_back({idx, val, lo_, hi_});
  }
};

struct IlpResult {
  enum Status { kOptimal, kFeasible /*time limit hit, incumbent returned*/, kInfeasible, kUnbounded } status = kInfeasible;
  std::vector<double> x;
  double objective = 0;
  long nodes = 0;
  double seconds = 0;
  std::string StatusName() const;
};

// LP relaxation only (continuous).
IlpResult SolveLp(const IlpModel& m);
// `cutoff`: only solutions strictly better than this objective are of interest (prunes the search).
IlpResult SolveIlp(const IlpModel& m, double time_limit_s = 60.0, double cutoff = IlpModel::kInf);

}  // namespace tepdist
