// ModelBase: thin caller of the C-ABI model stage kernels (no arithmetic on the host).
#include "model_base.hpp"

#include <cstdio>
#include <cstdlib>

#include "mppi_b200.h"

namespace {
void check(int rc, const char *what)
{
    if (rc != MPPI_OK) {   // the reference aborts through TF_CHECK_OK (src/model_base.cpp:43)
        std::fprintf(stderr, "ModelBase::%s failed: %s\n", what, mppi_last_error(nullptr));
        std::abort();
    }
}
}  // namespace

ModelBase::ModelBase() : m_dt(0.01f), m_m(1.f), m_s_dim(2), m_a_dim(1) {}
ModelBase::ModelBase(const float mass, const float dt, const int s_dim, const int a_dim)
    : m_dt(dt), m_m(mass), m_s_dim(s_dim), m_a_dim(a_dim) {}
ModelBase::~ModelBase() {}

std::vector<float> ModelBase::predict(const std::vector<float> &state, const std::vector<float> &action) const
{
    const int kst = (int)state.size() / m_s_dim, k = (int)action.size() / m_a_dim;
    std::vector<float> out((size_t)k * m_s_dim);
    check(mppi_model_step(m_device, m_m, m_dt, m_s_dim, m_a_dim, kst, k, state.data(), action.data(), out.data()),
          "predict");
    return out;
}

std::vector<float> ModelBase::freeStep(const std::vector<float> &state) const
{
    const int kst = (int)state.size() / m_s_dim;
    std::vector<float> out(state.size());
    check(mppi_model_free_step(m_device, m_m, m_dt, m_s_dim, m_a_dim, kst, state.data(), out.data()), "freeStep");
    return out;
}

std::vector<float> ModelBase::actionStep(const std::vector<float> &action) const
{
    const int k = (int)action.size() / m_a_dim;
    std::vector<float> out((size_t)k * m_s_dim);
    check(mppi_model_action_step(m_device, m_m, m_dt, m_s_dim, m_a_dim, k, action.data(), out.data()), "actionStep");
    return out;
}

std::vector<float> ModelBase::A() const
{
    const float blk[4] = {1.f, m_dt, 0.f, 1.f};
    std::vector<float> out((size_t)m_s_dim * m_s_dim);
    check(mppi_block_diag(blk, 2, 2, m_s_dim / 2, out.data()), "A");
    return out;
}

std::vector<float> ModelBase::B() const
{
    const float blk[2] = {(m_dt * m_dt) / 2.f / m_m, m_dt / m_m};
    std::vector<float> out((size_t)m_s_dim * m_a_dim);
    check(mppi_block_diag(blk, 2, 1, m_a_dim, out.data()), "B");
    return out;
}

void ModelBase::train() {}
