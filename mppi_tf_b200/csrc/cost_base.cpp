// CostBase: thin caller of the C-ABI cost stage kernels (no arithmetic on the host).
#include "cost_base.hpp"

#include <cstdio>
#include <cstdlib>

#include "mppi_b200.h"

namespace {
void check(int rc, const char *what)
{
    if (rc != MPPI_OK) {
        std::fprintf(stderr, "CostBase::%s failed: %s\n", what, mppi_last_error(nullptr));
        std::abort();
    }
}
}  // namespace

CostBase::CostBase() {}
CostBase::CostBase(const float lambda, const std::vector<float> sigma, const std::vector<float> goal)
    : in_lambda(lambda), in_sigma(sigma), m_goal(goal), in_Q(goal.size(), 1.f) {}
CostBase::CostBase(const float lambda, const std::vector<float> sigma, const std::vector<float> goal,
                   const std::vector<float> Q)
    : in_lambda(lambda), in_sigma(sigma), m_goal(goal), in_Q(Q) {}
CostBase::~CostBase() {}

bool CostBase::setGoal(std::vector<float> goal)
{
    if (goal.size() != m_goal.size()) return false;
    m_goal = std::move(goal);
    return true;
}

std::vector<float> CostBase::stateCost(const std::vector<float> &state) const
{
    const int s = (int)m_goal.size(), k = (int)state.size() / s;
    std::vector<float> out(k);
    check(mppi_cost_state(m_device, k, s, state.data(), m_goal.data(), in_Q.data(), out.data()), "stateCost");
    return out;
}

static int a_dim_of(const std::vector<float> &sigma)
{
    int a = 1;
    while ((size_t)a * a < sigma.size()) a++;
    return a;
}

std::vector<float> CostBase::actionCost(const std::vector<float> &action, const std::vector<float> &noise) const
{
    const int a = a_dim_of(in_sigma), k = (int)noise.size() / a;
    std::vector<float> out(k);
    check(mppi_cost_action(m_device, k, a, in_lambda, in_sigma.data(), action.data(), noise.data(), out.data()),
          "actionCost");
    return out;
}

std::vector<float> CostBase::stepCost(const std::vector<float> &state, const std::vector<float> &action,
                                      const std::vector<float> &noise) const
{
    const int s = (int)m_goal.size(), a = a_dim_of(in_sigma), k = (int)state.size() / s;
    std::vector<float> out(k);
    check(mppi_cost_step(m_device, k, s, a, in_lambda, in_sigma.data(), m_goal.data(), in_Q.data(), state.data(),
                         action.data(), noise.data(), out.data()),
          "stepCost");
    return out;
}
