// ControllerBase — the MPPI controller, B200-native.
// Same class name, constructor and numerical public methods as
// /root/reference/include/controller_base.hpp:16-376 (next, setGoal, saveNext, toCSV); the
// TensorFlow session is replaced by a handle of the C-ABI (include/mppi_b200.h) and the m*
// graph-builder methods by stage calls on plain buffers so that the unit KATs of
// test/test_controller.cpp can be restated 1:1 (tests/cpp/test_kats.cpp).
#ifndef MPPI_B200_CONTROLLER_BASE_HPP
#define MPPI_B200_CONTROLLER_BASE_HPP

#include <string>
#include <vector>

#include "cost_base.hpp"
#include "data_base.hpp"
#include "model_base.hpp"
#include "mppi_b200.h"

class ControllerBase {
public:
    ControllerBase();
    // controller_base.hpp:60-65.  As in the reference (src/controller_base.cpp:37-69): lambda = 1,
    // sigma = I, goal = (1,0,1,0,..), Q = ones, and the model is built with mass 1 — the `mass`
    // argument is stored but NOT forwarded to ModelBase (:68).  Use setModelMass to change it.
    ControllerBase(const int k, const int tau, const float dt, const float mass, const int s_dim,
                   const int a_dim);
    // B200 additions: explicit constants, sample sharding (rank/world) and batching.
    ControllerBase(const mppi_config &cfg);
    ~ControllerBase();
    ControllerBase(const ControllerBase &) = delete;
    ControllerBase &operator=(const ControllerBase &) = delete;

    void toCSV(std::string filename);                        // :93
    bool setGoal(std::vector<float> goal);                   // :95  (size mismatch -> cerr + false)
    void saveNext(std::vector<float> x_next);                // :97
    std::vector<float> next(std::vector<float> x);           // :109 (aborts on device failure, like TF_CHECK_OK)

    // parity / debug: the same update with eps [k][tau][a] injected instead of generated
    std::vector<float> nextWithNoise(std::vector<float> x, const std::vector<float> &eps);
    std::vector<float> getCosts();                           // [k] per-sample costs of the last update
    std::vector<float> getSequence();                        // m_U, [tau][a]
    std::vector<float> getUpdate();                          // U + sum w eps, pre-shift
    std::vector<float> dumpNoise();                          // eps of the last generated update
    void setSequence(const std::vector<float> &U);
    bool setLambda(float lambda);
    bool setSigma(const std::vector<float> &sigma);
    bool setQ(const std::vector<float> &q);
    bool setModelMass(float mass);
    // Python-twin extras (scripts/src/costs/cost_base.py:114-170, controllers/controller_base.py:368,468-474):
    // python_form = false keeps lambda u^T S^-1 eps (src/cost_base.cpp:63-68); upsilon scales the sampling
    bool setActionCost(bool python_form, float gamma, float upsilon);
    bool setNormalizeCost(bool on);

    // stage entry points (the reference's m* methods, :123-359) on plain buffers
    float mBeta(const std::vector<float> &cost);                                            // :123
    std::vector<float> mExpArg(const std::vector<float> &cost, float beta);                 // :137
    std::vector<float> mExp(const std::vector<float> &arg);                                 // :153
    float mNabla(const std::vector<float> &exp);                                            // :167
    std::vector<float> mWeights(const std::vector<float> &exp, float nabla);                // :182
    std::vector<float> mWeightedNoise(const std::vector<float> &weights, const std::vector<float> &noises);  // :199
    std::vector<float> mPrepareAction(const std::vector<float> &actions, int timestep);     // :217
    std::vector<float> mPrepareNoise(const std::vector<float> &noises, int timestep);       // :235
    std::vector<float> mShift(const std::vector<float> &current, const std::vector<float> &init, int nb);   // :270
    std::vector<float> mInit0(int nb);                                                      // :359
    std::vector<float> mGetNew(const std::vector<float> &current, int nb);                  // :290

    mppi_handle *handle() { return m_h; }
    const char *lastError() const;

private:
    void create(const mppi_config &cfg);
    void die(const char *what) const;

    float m_dt = 0, m_mass = 0;
    int m_k = 0, m_tau = 0, m_s_dim = 0, m_a_dim = 0;
    float m_lambda = 1.f;
    DataBase m_db;
    mppi_handle *m_h = nullptr;
};

#endif
