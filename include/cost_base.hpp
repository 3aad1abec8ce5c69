// CostBase — quadratic state cost + lambda u^T Sigma^-1 eps action cost on the GPU.
// Same class name and constructor meaning as /root/reference/include/cost_base.hpp:7-178
// (tensors passed as std::vector<float>); the graph builders become batched calls through the
// C-ABI.  Unlike the reference (cost_base.hpp:98-100: no return, goal baked into the graph),
// setGoal takes effect and returns true.
#ifndef MPPI_B200_COST_BASE_HPP
#define MPPI_B200_COST_BASE_HPP

#include <vector>

class CostBase {
public:
    CostBase();
    // (lambda, sigma [a][a], goal [s]) — cost_base.hpp:47-49 (declared, never defined there); Q = ones
    CostBase(const float lambda, const std::vector<float> sigma, const std::vector<float> goal);
    // (lambda, sigma [a][a], goal [s], Q [s] diagonal) — cost_base.hpp:77-80
    CostBase(const float lambda, const std::vector<float> sigma, const std::vector<float> goal,
             const std::vector<float> Q);
    ~CostBase();

    bool setGoal(std::vector<float> goal);
    // mStateCost (:153-154) / mBuildFinalStepCostGraph (:137-138): state [k][s] -> [k]
    std::vector<float> stateCost(const std::vector<float> &state) const;
    std::vector<float> finalCost(const std::vector<float> &state) const { return stateCost(state); }
    // mActionCost (:167-169): action [a] (the un-perturbed U[t]), noise [k][a] -> [k]
    std::vector<float> actionCost(const std::vector<float> &action, const std::vector<float> &noise) const;
    // mBuildStepCostGraph (:119-122)
    std::vector<float> stepCost(const std::vector<float> &state, const std::vector<float> &action,
                                const std::vector<float> &noise) const;

    float lambda() const { return in_lambda; }
    const std::vector<float> &sigma() const { return in_sigma; }
    const std::vector<float> &goal() const { return m_goal; }
    const std::vector<float> &Q() const { return in_Q; }
    void setDevice(int device) { m_device = device; }

private:
    float in_lambda = 1.f;
    std::vector<float> in_sigma, m_goal, in_Q;
    int m_device = -1;
};

#endif
