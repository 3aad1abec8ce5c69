// The reference's gtest suites for the hot path (test/test_{utile,model,cost,controller}.cpp under
// /root/reference) restated against the B200-native C++ classes.  gtest is not available offline,
// so a ~20-line harness stands in; EXPECT_FLOAT_EQ keeps gtest's meaning (within 4 ULPs).
// Built by mppi_tf_b200/csrc/Makefile, run by tests/test_cpp_host.py on the GPU box.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "controller_base.hpp"
#include "cost_base.hpp"
#include "model_base.hpp"

using std::vector;

static int g_fail = 0, g_checks = 0;

static bool almost_equal_4ulp(float a, float b)
{
    if (std::isnan(a) || std::isnan(b)) return false;
    int32_t ia, ib;
    std::memcpy(&ia, &a, 4);
    std::memcpy(&ib, &b, 4);
    auto biased = [](int32_t i) { return (uint32_t)(i < 0 ? ~i + 1 : i | 0x80000000u); };   // gtest's SignAndMagnitudeToBiased
    const uint32_t ua = biased(ia), ub = biased(ib);
    return (ua > ub ? ua - ub : ub - ua) <= 4;
}

static void expect_vec(const vector<float> &got, const vector<float> &want, const char *name)
{
    g_checks++;
    if (got.size() != want.size()) {
        std::printf("FAIL %s: size %zu != %zu\n", name, got.size(), want.size());
        g_fail++;
        return;
    }
    for (size_t i = 0; i < got.size(); i++)
        if (!almost_equal_4ulp(got[i], want[i])) {
            std::printf("FAIL %s[%zu]: %.9g != %.9g\n", name, i, got[i], want[i]);
            g_fail++;
            return;
        }
}
static void expect_true(bool c, const char *name)
{
    g_checks++;
    if (!c) { std::printf("FAIL %s\n", name); g_fail++; }
}

// ---- test/test_model.cpp ---------------------------------------------------------------------------
static void model_tests()
{
    {   // StepTesting1 :120-146
        const float dt = 0.01f, m = 1.f;
        ModelBase model(m, dt, 2, 1);
        const float acc = (dt * dt) / (2.f * m), vel = dt / m;
        expect_vec(model.freeStep({0.f, 0.f}), {0.f, 0.f}, "StepTesting1.free");
        expect_vec(model.actionStep({1.f}), {acc, vel}, "StepTesting1.action");
        expect_vec(model.predict({0.f, 0.f}, {1.f}), {acc, vel}, "StepTesting1.result");
    }
    {   // StepTesting2 :148-176
        const float dt = 0.01f, m = 2.f;
        ModelBase model(m, dt, 4, 2);
        const float acc = (dt * dt) / (2.f * m), vel = dt / m;
        expect_vec(model.actionStep({1.f, 1.f}), {acc, vel, acc, vel}, "StepTesting2.action");
        expect_vec(model.predict({0, 0, 0, 0}, {1.f, 1.f}), {acc, vel, acc, vel}, "StepTesting2.result");
    }
    const float dt3 = 0.01f, m3 = 1.5f;
    const float acc = (dt3 * dt3) / (2.f * m3), vel = dt3 / m3;
    const vector<float> state3 = {0., 0., 0., 0., 0., 0., 2., 1., 5., 0., -1., -2., 0.5, 0.5, 0.5, 0.5, 0.5, 0.5,
                                  1., 0., 1., 0., 1., 0., -1, 0.5, -3, 2., 0., 0.};
    const vector<float> action3 = {1., 1., 1., 2., 0., -1., 0., 0., 0., 0.5, -0.5, 0.5, 3., 3., 3.};
    const vector<float> exp_u = {acc, vel, acc, vel, acc, vel,
                                 2.f * acc, 2.f * vel, 0.f * acc, 0 * vel, -1.f * acc, -1.f * vel,
                                 0.f * acc, 0.f * vel, 0.f * acc, 0 * vel, 0.f * acc, 0.f * vel,
                                 0.5f * acc, 0.5f * vel, -0.5f * acc, -0.5f * vel, 0.5f * acc, 0.5f * vel,
                                 3.f * acc, 3.f * vel, 3.f * acc, 3.f * vel, 3.f * acc, 3.f * vel};
    ModelBase model3(m3, dt3, 6, 3);
    {   // LargeTesting :178-214
        const vector<float> exp_s = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f,
                                     2.f + dt3, 1.f, 5.f, 0.f, -1.f - 2.f * dt3, -2.f,
                                     0.5f + dt3 / 2.f, 0.5f, 0.5f + dt3 / 2.f, 0.5f, 0.5f + dt3 / 2.f, 0.5f,
                                     1.f, 0.f, 1.f, 0.f, 1.f, 0.f,
                                     -1.f + dt3 / 2.f, 0.5f, -3.f + 2.f * dt3, 2.f, 0.f, 0.f};
        vector<float> exp_res;
        for (size_t i = 0; i < exp_u.size(); i++) exp_res.push_back(exp_u[i] + exp_s[i]);
        expect_vec(model3.freeStep(state3), exp_s, "LargeTesting.free");
        expect_vec(model3.actionStep(action3), exp_u, "LargeTesting.action");
        expect_vec(model3.predict(state3, action3), exp_res, "LargeTesting.result");
    }
    {   // InitTest :216-255 — [1,s,1] state broadcast over [k,a,1] actions
        const vector<float> init = {-1, 0.5, -3, 2., 0., 0.};
        const vector<float> row = {-1.f + dt3 / 2.f, 0.5f, -3.f + 2.f * dt3, 2.f, 0.f, 0.f};
        vector<float> exp_res;
        for (int k = 0; k < 5; k++)
            for (int j = 0; j < 6; j++) exp_res.push_back(exp_u[k * 6 + j] + row[j]);
        expect_vec(model3.freeStep(init), row, "InitTest.free");
        expect_vec(model3.predict(init, action3), exp_res, "InitTest.result");
    }
    {   // test/test_utile.cpp:83-108 blockDiagTest2 through ModelBase's A and B (dt=.01, m=1.5)
        ModelBase m2(1.5f, 0.01f, 4, 2);
        const float dt = 0.01f, m = 1.5f;
        expect_vec(m2.A(), {1.f, dt, 0, 0, 0, 1.f, 0, 0, 0, 0, 1.f, dt, 0, 0, 0, 1.f}, "blockDiag2.A");
        expect_vec(m2.B(), {(dt * dt) / (2 * m), 0, dt / m, 0, 0, (dt * dt) / (2 * m), 0, dt / m}, "blockDiag2.B");
    }
}

// ---- test/test_cost.cpp --------------------------------------------------------------------------------
static void cost_tests()
{
    CostBase c1(1.f, {1., 0., 0., 1.}, {1., 1.}, {1., 1.});                        // :29-54
    expect_vec(c1.finalCost({0., 1.}), {1.}, "StateCost.1");                        // :172-178
    expect_vec(c1.stepCost({0., 1.}, {1., 1.}, {1., 1.}), {3.}, "StepCost.1");      // :207-215
    CostBase c2(1.f, {1., 0., 0., 1.}, {1., 1., 1., 2.}, {1., 1., 10., 10.});       // :56-84
    expect_vec(c2.finalCost({0., 0.5, 2., 0.}), {51.25}, "StateCost.2");            // :181-189
    expect_vec(c2.stepCost({0., 0.5, 2., 0.}, {0.5, 2.}, {0.5, 1.}), {53.5}, "StepCost.2");   // :218-225
    CostBase c3(1.f, {1., 0., 0., 0., 1., 0., 0., 0., 1.}, {1., 1., 1., 2.}, {1., 1., 10., 10.});   // :86-127
    const vector<float> state3 = {0., 0.5, 2., 0., 0., 2., 0., 0., 10., 2., 2., 3, 1., 1., 1., 2., 3., 4., 5., 6.};
    const vector<float> eps3 = {0.5, 1., 2., 0.5, 2., 0.25, -2, -0.2, -1, 0, 0, 0, 1., 0.5, 3.};
    expect_vec(c3.finalCost(state3), {51.25, 52, 102, 0., 333}, "StateCost.3");     // :193-201
    expect_vec(c3.stepCost(state3, {0.5, 2., 0.25}, eps3),
               {51.25f + 2.75f, 52.f + 4.3125f, 102.f - 1.65f, 0.f, 333.f + 2.25f}, "StepCost.3");   // :228-237
    expect_true(c2.setGoal({0., 0.5, 2., 0.}), "CostBase.setGoal returns true");
    expect_vec(c2.stateCost({0., 0.5, 2., 0.}), {0.}, "CostBase.setGoal takes effect");
    expect_true(!c2.setGoal({1., 2.}), "CostBase.setGoal size mismatch");
}

// ---- test/test_controller.cpp ----------------------------------------------------------------------------
static void controller_tests()
{
    const int k = 5, tau = 3, a_dim = 2;
    ControllerBase cont(k, tau, 0.01f, 1.f, 4, a_dim);                              // :17-19
    const vector<float> cost = {3., 10., 0., 1., 5.};
    const vector<float> noise = {1., -0.5, 1., -0.5, 2., 1., 0.3, 0, 2., 0.2, 1.2, 3., 0.5, 0.5, 0.5, 0.5, 0.5, 0.5,
                                 0.6, 0.7, 0.2, -0.3, 0.1, -0.4, -2., -3., -4., -1., 0., 0.};
    const vector<float> action = {1., 0.5, 2.3, 4.5, 2.1, -0.4};
    // testDataPrep :71-107
    expect_vec(cont.mPrepareAction(action, 0), {1., 0.5}, "DataPrep.a0");
    expect_vec(cont.mPrepareAction(action, 1), {2.3, 4.5}, "DataPrep.a1");
    expect_vec(cont.mPrepareAction(action, 2), {2.1, -0.4}, "DataPrep.a2");
    expect_vec(cont.mPrepareNoise(noise, 0), {1., -0.5, 0.3, 0, 0.5, 0.5, 0.6, 0.7, -2., -3.}, "DataPrep.n0");
    expect_vec(cont.mPrepareNoise(noise, 1), {1., -0.5, 2., 0.2, 0.5, 0.5, 0.2, -0.3, -4, -1}, "DataPrep.n1");
    expect_vec(cont.mPrepareNoise(noise, 2), {2., 1., 1.2, 3., 0.5, 0.5, 0.1, -0.4, 0., 0.}, "DataPrep.n2");
    // testUpdate :109-167, chained exactly like the reference test
    const float b = cont.mBeta(cost);
    const vector<float> e_arg = cont.mExpArg(cost, b);
    const vector<float> e = cont.mExp(e_arg);
    const float nab = cont.mNabla(e);
    const vector<float> w = cont.mWeights(e, nab);
    const vector<float> w_n = cont.mWeightedNoise(w, noise);
    expect_vec({b}, {0.f}, "Update.beta");
    expect_vec(e_arg, {-3., -10., 0, -1., -5.}, "Update.exp_arg");
    expect_vec(e, {0.049787068367863944, 4.5399929762484854e-05, 1, 0.36787944117144233, 0.006737946999085467},
               "Update.exp");
    expect_vec({nab}, {1.424449856468154}, "Update.nabla");
    const double W[5] = {0.034951787275480706, 3.1871904480408675e-05, 0.7020254138530686, 0.2582607169364174,
                         0.004730210030553017};
    expect_vec(w, {(float)W[0], (float)W[1], (float)W[2], (float)W[3], (float)W[4]}, "Update.weights");
    expect_vec(w_n,
               {(float)(W[0] * 1. + W[1] * 0.3 + W[2] * 0.5 + W[3] * 0.6 + W[4] * (-2)),
                (float)(W[0] * (-0.5) + W[1] * 0 + W[2] * 0.5 + W[3] * 0.7 + W[4] * (-3)),
                (float)(W[0] * 1 + W[1] * 2 + W[2] * 0.5 + W[3] * 0.2 + W[4] * (-4)),
                (float)(W[0] * (-0.5) + W[1] * 0.2 + W[2] * 0.5 + W[3] * (-0.3) + W[4] * (-1)),
                (float)(W[0] * 2 + W[1] * 1.2 + W[2] * 0.5 + W[3] * 0.1 + W[4] * 0),
                (float)(W[0] * 1 + W[1] * 3 + W[2] * 0.5 + W[3] * (-0.4) + W[4] * 0)},
               "Update.weighted_noise");
    float sum_w = 0.f;
    for (float v : w) sum_w += v;
    expect_vec({sum_w}, {1.f}, "Update.sum_w");
    // testNew :169-193
    expect_vec(cont.mGetNew(action, 0), {}, "New.0");
    expect_vec(cont.mGetNew(action, 1), {1, 0.5}, "New.1");
    expect_vec(cont.mGetNew(action, 2), {1, 0.5, 2.3, 4.5}, "New.2");
    expect_vec(cont.mGetNew(action, 3), {1., 0.5, 2.3, 4.5, 2.1, -0.4}, "New.3");
    // testShiftAndInit :195-222
    expect_vec(cont.mShift(action, {1, 0.5}, 1), {2.3, 4.5, 2.1, -0.4, 1., 0.5}, "Shift.1");
    expect_vec(cont.mShift(action, {1, 0.5, 2.3, 4.5}, 2), {2.1, -0.4, 1., 0.5, 2.3, 4.5}, "Shift.2");
    expect_vec(cont.mInit0(2), {0, 0, 0, 0}, "Init0");
    // setGoal :126-133
    expect_true(!cont.setGoal({1., 1., 1.}), "setGoal size mismatch -> false");
    expect_true(cont.setGoal({1., 1., 1., 2.}), "setGoal ok");
}

// ---- the caller contract of src/main.cpp:36-45 (commented loop): next / saveNext / toCSV ----------------------
static void closed_loop_test()
{
    const int k = 2048, tau = 20, s = 2, a = 1;
    ControllerBase ctrl(k, tau, 0.1f, 1.f, s, a);
    ModelBase plant(1.f, 0.1f, s, a);
    vector<float> x = {0.f, 0.f};
    float first_dist = 0, last_dist = 0;
    for (int i = 0; i < 40; i++) {
        vector<float> u = ctrl.next(x);
        expect_true((int)u.size() == a && std::isfinite(u[0]), "next returns a finite action");
        x = plant.predict(x, u);
        ctrl.saveNext(x);
        const float d = std::fabs(x[0] - 1.f);     // default goal (1, 0)
        if (i == 0) first_dist = d;
        last_dist = d;
    }
    expect_true(last_dist < 0.5f * first_dist, "closed loop moves the point mass toward the goal");
    // injected-noise parity entry point: zero noise leaves the sequence unchanged before the shift
    vector<float> U = ctrl.getSequence();
    vector<float> zeros((size_t)k * tau * a, 0.f);
    ctrl.nextWithNoise(x, zeros);
    vector<float> upd = ctrl.getUpdate();
    expect_vec(upd, U, "zero noise: U' == U");
    vector<float> costs = ctrl.getCosts();
    expect_true(costs.size() == (size_t)k && costs[0] == costs[k - 1], "zero noise: identical costs");
    const char *path = "/tmp/mppi_b200_test.csv";
    ctrl.toCSV(path);
    std::ifstream f(path);
    std::string header, row;
    std::getline(f, header);
    std::getline(f, row);
    expect_true(header == "x0,x1,u0,x_next0,x_next1,", "CSV header format (src/data_base.cpp:43-50,62-64)");
    expect_true(std::count(row.begin(), row.end(), ',') == 5 && row.back() == ',', "CSV row format");
}

int main()
{
    model_tests();
    cost_tests();
    controller_tests();
    closed_loop_test();
    std::printf("%d checks, %d failures\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
