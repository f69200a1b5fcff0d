// Drives the MEX gateway the way src/TinyMPC.m does (string command + mxArrays) and prints results as JSON-ish
// lines that tests/test_mex_gateway.py parses.  Usage: mex_driver host|gpu
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mex.h"

static mxArray* M(int r, int c, std::initializer_list<double> rowmajor) {
    mxArray* a = mxCreateDoubleMatrix(r, c, mxREAL);
    int k = 0;
    for (double v : rowmajor) { int i = k / c, j = k % c; mxGetPr(a)[j * r + i] = v; ++k; }
    return a;
}
static mxArray* S(double v) { return mxCreateDoubleScalar(v); }
static mxArray* F(int r, int c, double v) { mxArray* a = mxCreateDoubleMatrix(r, c, mxREAL); for (int i = 0; i < r * c; ++i) mxGetPr(a)[i] = v; return a; }

static std::vector<mxArray*> call(const char* cmd, std::vector<mxArray*> args, int nlhs = 1) {
    std::vector<const mxArray*> in{mxCreateString(cmd)};
    for (auto* a : args) in.push_back(a);
    std::vector<mxArray*> out(8, nullptr);
    mexFunction(nlhs, out.data(), (int)in.size(), in.data());
    return out;
}

int main(int argc, char** argv) {
    const bool gpu = argc > 1 && std::string(argv[1]) == "gpu";
    // examples/cartpole_example_one_solve.m:13-31 + set_bound_constraints([],[],-0.5,0.5) as TinyMPC.m expands it
    mxArray* A = M(4, 4, {1, 0.01, 0, 0, 0, 1, 0.039, 0, 0, 0, 1.002, 0.01, 0, 0, 0.458, 1.002});
    mxArray* B = M(4, 1, {0, 0.02, 0, 0.067});
    mxArray* Q = M(4, 4, {10, 0, 0, 0, 0, 1, 0, 0, 0, 0, 10, 0, 0, 0, 0, 1});
    mxArray* R = M(1, 1, {1});
    auto o = call("setup", {A, B, F(4, 1, 0), Q, R, S(1.0), S(4), S(1), S(20), S(0)});
    std::printf("setup_status %g\n", mxGetScalar(o[0]));
    // TinyMPC.m pushes its own defaults right after setup (src/TinyMPC.m:94-98)
    call("update_settings", {S(1e-4), S(1e-4), S(100), S(1), S(0), S(0), S(0), S(0), S(0), S(0), S(0), S(0.1), S(10), S(1), S(0)}, 0);
    try { call("bogus", {}, 0); std::printf("bogus_ok\n"); } catch (const MexError& e) { std::printf("error_id %s\n", e.id.c_str()); }
    try { call("set_x0", {}, 0); } catch (const MexError& e) { std::printf("error_id %s\n", e.id.c_str()); }
    call("set_bound_constraints", {F(4, 20, -1e17), F(4, 20, 1e17), F(1, 19, -0.5), F(1, 19, 0.5), S(0)}, 0);
    call("update_settings", {S(1e-4), S(1e-4), S(100), S(1), S(1), S(1), S(0), S(0), S(0), S(0), S(0), S(0.1), S(10), S(1), S(0)}, 0);
    call("set_x0", {M(4, 1, {0.5, 0, 0, 0}), S(0)}, 0);
    if (!gpu) { std::printf("host_only_done\n"); return 0; }
    o = call("solve", {S(0)});
    std::printf("solve_ret %g\n", mxGetScalar(o[0]));
    o = call("get_stats", {S(0)}, 4);
    std::printf("iter %g status %g\n", mxGetScalar(o[0]), mxGetScalar(o[1]));
    o = call("get_solution", {S(0)}, 2);
    std::printf("u");
    for (int i = 0; i < 19; ++i) std::printf(" %.10f", mxGetPr(o[1])[i]);
    std::printf("\n");
    // closed loop on the ONE solver of the gateway (examples/cartpole_example_mpc.m:36-44): x0 <- A x0 + B u_sol(1), solve again
    double x0n[4];
    {
        const double x0[4] = {0.5, 0, 0, 0}, u0 = mxGetPr(o[1])[0];
        for (int i = 0; i < 4; ++i) {
            x0n[i] = mxGetPr(B)[i] * u0;
            for (int j = 0; j < 4; ++j) x0n[i] += mxGetPr(A)[j * 4 + i] * x0[j];
        }
        call("set_x0", {M(4, 1, {x0n[0], x0n[1], x0n[2], x0n[3]}), S(0)}, 0);
        call("solve", {S(0)});
        auto st2 = call("get_stats", {S(0)}, 4);
        std::printf("loop_iter2 %g\n", mxGetScalar(st2[0]));
    }
    // solve_batch: 3 copies of the same x0 (double input) -> every problem must stop at the same iteration
    mxArray* X0 = mxCreateDoubleMatrix(4, 3, mxREAL);
    for (int b = 0; b < 3; ++b) mxGetPr(X0)[4 * b] = 0.5;
    mxArray* E = mxCreateDoubleMatrix(0, 0, mxREAL);
    call("set_option", {mxCreateString("precision"), S(64)}, 0);
    o = call("solve_batch", {X0, E, E, E, E, E, E, S(0)}, 4);
    const int* it = static_cast<const int*>(mxGetData(o[2]));
    const float* U = static_cast<const float*>(mxGetData(o[1]));
    std::printf("batch_iter %d %d %d dims %zu u0 %.7f\n", it[0], it[1], it[2], mxGetNumberOfDimensions(o[0]), U[0]);
    // sessions: 3 warm-started copies of the solver on the GPU, the same closed loop (precision 64 was set above)
    try { call("session_solve", {}, 0); } catch (const MexError& e) { std::printf("session_error_id %s\n", e.id.c_str()); }
    call("session_create", {S(3)}, 0);
    call("session_set_x0", {X0}, 0);
    call("session_solve", {}, 0);
    o = call("session_read", {mxCreateString("iter")});
    std::printf("session_iter %g %g %g\n", mxGetPr(o[0])[0], mxGetPr(o[0])[1], mxGetPr(o[0])[2]);
    call("session_step", {S(1)}, 0);
    o = call("session_read", {mxCreateString("x0")});
    double dmax = 0;
    for (int b = 0; b < 3; ++b) for (int i = 0; i < 4; ++i) { const double d = mxGetPr(o[0])[4 * b + i] - x0n[i]; dmax = d > dmax ? d : (-d > dmax ? -d : dmax); }
    std::printf("session_x0_err %.3e\n", dmax);
    call("session_solve", {}, 0);
    o = call("session_read", {mxCreateString("iter")});
    std::printf("session_iter2 %g %g %g\n", mxGetPr(o[0])[0], mxGetPr(o[0])[1], mxGetPr(o[0])[2]);
    o = call("session_read", {mxCreateString("sol_u")});
    std::printf("session_dims %zu\n", mxGetNumberOfDimensions(o[0]));
    try { call("session_read", {mxCreateString("bogus")}); } catch (const MexError& e) { std::printf("session_error_id %s\n", e.id.c_str()); }
    call("session_destroy", {}, 0);
    call("reset", {S(0)}, 0);
    try { call("solve", {S(0)}); } catch (const MexError& e) { std::printf("error_id %s\n", e.id.c_str()); }
    return 0;
}
