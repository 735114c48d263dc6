#include "b200_solvers.hpp"
namespace sqaod { namespace cuda {
template <> BipartiteGraphAnnealer<float> *newBipartiteGraphAnnealer<float>() { sqb_throwError("not implemented"); return 0; }
template <> BipartiteGraphAnnealer<double> *newBipartiteGraphAnnealer<double>() { sqb_throwError("not implemented"); return 0; }
template <> DenseGraphBFSearcher<float> *newDenseGraphBFSearcher<float>() { sqb_throwError("not implemented"); return 0; }
template <> DenseGraphBFSearcher<double> *newDenseGraphBFSearcher<double>() { sqb_throwError("not implemented"); return 0; }
template <> BipartiteGraphBFSearcher<float> *newBipartiteGraphBFSearcher<float>() { sqb_throwError("not implemented"); return 0; }
template <> BipartiteGraphBFSearcher<double> *newBipartiteGraphBFSearcher<double>() { sqb_throwError("not implemented"); return 0; }
}}
