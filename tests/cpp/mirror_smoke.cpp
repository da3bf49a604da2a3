// Compiles and links the C++ mirror against libfamg.so; run on a GPU it solves a 1-D Poisson
// problem with a two-level multigrid, without a GPU it must fail loudly (no CPU fallback).
#include <cstdio>
#include <vector>

#include "famg.hpp"

int main() {
    try {
        famg::Context ctx(0);
        const int64_t n = 63, nc = 31;
        std::vector<uint64_t> r, c; std::vector<double> v;
        for (int64_t i = 0; i < n; ++i) {
            if (i > 0) { r.push_back(i); c.push_back(i - 1); v.push_back(-1.0); }
            r.push_back(i); c.push_back(i); v.push_back(2.0);
            if (i + 1 < n) { r.push_back(i); c.push_back(i + 1); v.push_back(-1.0); }
        }
        auto a = famg::SparseRowMat::try_new_from_triplets(ctx, n, n, r, c, v);
        std::vector<uint64_t> pr, pc; std::vector<double> pv;
        for (int64_t j = 0; j < nc; ++j) for (int k = 0; k < 3; ++k) { pr.push_back(2 * j + k); pc.push_back(j); pv.push_back(k == 1 ? 1.0 : 0.5); }
        auto p = famg::SparseRowMat::try_new_from_triplets(ctx, n, nc, pr, pc, pv);
        auto rt = p->transpose();
        auto ac = (*rt) * *((*a) * (*p));
        famg::Multigrid mg(*a, *famg::new_jacobi(*a, 0.66));
        mg.add_level(*ac, *famg::SparseCholeskySolve(*ac), *rt, *p);
        std::vector<double> b(n, 1.0), x(n, 0.0);
        famg::CgParams prm; prm.rel_tolerance = 1e-10;
        famg_cg_info info = famg::conjugate_gradient(x.data(), mg, *a, b.data(), prm);
        std::printf("mirror ok: %lld iterations, rel residual %.2e\n", (long long)info.iter_count, info.rel_residual);
        return info.rel_residual < 1e-10 ? 0 : 1;
    } catch (const famg::Error &e) {
        std::printf("famg::Error %d: %s\n", e.status, e.what());
        return e.status == FAMG_ERR_CUDA ? 3 : 1;
    }
}
