// ControllerBase: the reference's public class over the C-ABI handle.  next() is one call into
// the fused CUDA update; everything numerical runs in libmppi_b200's kernels.
#include "controller_base.hpp"

#include <cstdio>
#include <cstdlib>
#include <iostream>

ControllerBase::ControllerBase() {}

ControllerBase::ControllerBase(const int k, const int tau, const float dt, const float mass, const int s_dim,
                               const int a_dim)
    : m_dt(dt), m_mass(mass), m_k(k), m_tau(tau), m_s_dim(s_dim), m_a_dim(a_dim)
{
    mppi_config cfg;
    // reference: m_model = ModelBase(1., m_dt, ...) — `mass` is kept in m_mass only
    // (/root/reference/src/controller_base.cpp:32,68)
    mppi_config_default(&cfg, k, tau, dt, 1.0f, s_dim, a_dim);
    create(cfg);
}

ControllerBase::ControllerBase(const mppi_config &cfg)
    : m_dt(cfg.dt), m_mass(cfg.mass), m_k(cfg.k), m_tau(cfg.tau), m_s_dim(cfg.s_dim), m_a_dim(cfg.a_dim)
{
    create(cfg);
}

void ControllerBase::create(const mppi_config &cfg)
{
    m_lambda = cfg.lambda;
    if (mppi_create(&cfg, &m_h) != MPPI_OK) {
        std::fprintf(stderr, "ControllerBase: %s\n", mppi_last_error(nullptr));
        std::abort();   // the reference aborts on graph/session failure (TF_CHECK_OK)
    }
}

ControllerBase::~ControllerBase() { mppi_destroy(m_h); }

const char *ControllerBase::lastError() const { return mppi_last_error(m_h); }

void ControllerBase::die(const char *what) const
{
    std::fprintf(stderr, "ControllerBase::%s failed: %s\n", what, mppi_last_error(m_h));
    std::abort();
}

bool ControllerBase::setGoal(std::vector<float> goal)
{
    if ((int)goal.size() != m_s_dim) {
        std::cerr << "Wrong goal size, it should match the state dimension: " << m_s_dim << std::endl;
        return false;
    }
    return mppi_set_goal_n(m_h, goal.data(), 1) == MPPI_OK;     // every controller of the handle gets this goal
}

std::vector<float> ControllerBase::next(std::vector<float> x)
{
    x.resize(m_s_dim);
    std::vector<float> act(m_a_dim);
    if (mppi_next(m_h, x.data(), act.data()) != MPPI_OK) die("next");
    m_db.addX(x);          // src/controller_base.cpp:146-147
    m_db.addU(act);
    return act;
}

std::vector<float> ControllerBase::nextWithNoise(std::vector<float> x, const std::vector<float> &eps)
{
    x.resize(m_s_dim);
    std::vector<float> act(m_a_dim);
    if (eps.size() != (size_t)mppi_k_local(m_h) * m_tau * m_a_dim) {
        std::cerr << "Wrong noise size, expected [k][tau][a_dim]" << std::endl;
        return {};
    }
    if (mppi_next_with_noise(m_h, x.data(), eps.data(), act.data()) != MPPI_OK) die("nextWithNoise");
    m_db.addX(x);
    m_db.addU(act);
    return act;
}

void ControllerBase::toCSV(std::string filename) { m_db.toCSV(filename); }

void ControllerBase::saveNext(std::vector<float> x_next)
{
    x_next.resize(m_s_dim);
    m_db.addNext(x_next);
}

std::vector<float> ControllerBase::getCosts()
{
    std::vector<float> out((size_t)mppi_k_local(m_h));
    if (mppi_get_costs(m_h, out.data()) != MPPI_OK) die("getCosts");
    return out;
}
std::vector<float> ControllerBase::getSequence()
{
    std::vector<float> out((size_t)m_tau * m_a_dim);
    if (mppi_get_sequence(m_h, out.data()) != MPPI_OK) die("getSequence");
    return out;
}
std::vector<float> ControllerBase::getUpdate()
{
    std::vector<float> out((size_t)m_tau * m_a_dim);
    if (mppi_get_update(m_h, out.data()) != MPPI_OK) die("getUpdate");
    return out;
}
std::vector<float> ControllerBase::dumpNoise()
{
    std::vector<float> out((size_t)mppi_k_local(m_h) * m_tau * m_a_dim);
    if (mppi_dump_noise(m_h, out.data()) != MPPI_OK) die("dumpNoise");
    return out;
}
void ControllerBase::setSequence(const std::vector<float> &U)
{
    if (U.size() != (size_t)m_tau * m_a_dim || mppi_set_sequence(m_h, U.data()) != MPPI_OK) die("setSequence");
}
bool ControllerBase::setLambda(float lambda)
{
    if (mppi_set_lambda(m_h, lambda) != MPPI_OK) return false;
    m_lambda = lambda;
    return true;
}
bool ControllerBase::setActionCost(bool python_form, float gamma, float upsilon)
{
    return mppi_set_action_cost(m_h, python_form ? MPPI_ACTION_COST_PYTHON : MPPI_ACTION_COST_CPP, gamma, upsilon) == MPPI_OK;
}

bool ControllerBase::setNormalizeCost(bool on) { return mppi_set_normalize_cost(m_h, on ? 1 : 0) == MPPI_OK; }

bool ControllerBase::setSigma(const std::vector<float> &sigma)
{
    return sigma.size() == (size_t)m_a_dim * m_a_dim && mppi_set_sigma(m_h, sigma.data()) == MPPI_OK;
}
bool ControllerBase::setQ(const std::vector<float> &q)
{
    return (int)q.size() == m_s_dim && mppi_set_q(m_h, q.data()) == MPPI_OK;
}
bool ControllerBase::setModelMass(float mass) { return mppi_set_mass(m_h, mass) == MPPI_OK; }

// ---- stage entry points: src/controller_base.cpp:166-213,310-329 -----------------------------------------
namespace {
std::vector<float> vector_op(int op, const std::vector<float> &in, float s0, float s1)
{
    const bool reduce = (op == MPPI_OP_MIN || op == MPPI_OP_SUM);
    std::vector<float> out(reduce ? 1 : in.size());
    if (mppi_stage_vector_op(-1, op, (int)in.size(), in.data(), s0, s1, out.data()) != MPPI_OK) {
        std::fprintf(stderr, "ControllerBase stage failed: %s\n", mppi_last_error(nullptr));
        std::abort();
    }
    return out;
}
}  // namespace

float ControllerBase::mBeta(const std::vector<float> &cost) { return vector_op(MPPI_OP_MIN, cost, 0.f, 1.f)[0]; }

std::vector<float> ControllerBase::mExpArg(const std::vector<float> &cost, float beta)
{
    return vector_op(MPPI_OP_EXP_ARG, cost, beta, m_lambda);
}

std::vector<float> ControllerBase::mExp(const std::vector<float> &arg) { return vector_op(MPPI_OP_EXP, arg, 0.f, 1.f); }

float ControllerBase::mNabla(const std::vector<float> &exp) { return vector_op(MPPI_OP_SUM, exp, 0.f, 1.f)[0]; }

std::vector<float> ControllerBase::mWeights(const std::vector<float> &exp, float nabla)
{
    return vector_op(MPPI_OP_DIV, exp, nabla, 1.f);
}

std::vector<float> ControllerBase::mWeightedNoise(const std::vector<float> &weights, const std::vector<float> &noises)
{
    const int k = (int)weights.size();
    const int TA = (int)(noises.size() / (size_t)k);
    std::vector<float> out(TA);
    if (mppi_weighted_noise(-1, k, TA, weights.data(), noises.data(), out.data()) != MPPI_OK) {
        std::fprintf(stderr, "ControllerBase::mWeightedNoise failed: %s\n", mppi_last_error(nullptr));
        std::abort();
    }
    return out;
}

std::vector<float> ControllerBase::mPrepareAction(const std::vector<float> &actions, int timestep)
{
    std::vector<float> out(m_a_dim);
    if (mppi_prepare_action((int)actions.size() / m_a_dim, m_a_dim, actions.data(), timestep, out.data()) != MPPI_OK)
        die("mPrepareAction");
    return out;
}

std::vector<float> ControllerBase::mPrepareNoise(const std::vector<float> &noises, int timestep)
{
    const int k = (int)(noises.size() / ((size_t)m_tau * m_a_dim));
    std::vector<float> out((size_t)k * m_a_dim);
    if (mppi_prepare_noise(-1, k, m_tau, m_a_dim, noises.data(), timestep, out.data()) != MPPI_OK) die("mPrepareNoise");
    return out;
}

std::vector<float> ControllerBase::mShift(const std::vector<float> &current, const std::vector<float> &init, int nb)
{
    std::vector<float> out(current.size());
    if (mppi_shift((int)current.size() / m_a_dim, m_a_dim, current.data(), init.data(), nb, out.data()) != MPPI_OK)
        die("mShift");
    return out;
}

std::vector<float> ControllerBase::mInit0(int nb) { return std::vector<float>((size_t)nb * m_a_dim, 0.f); }

std::vector<float> ControllerBase::mGetNew(const std::vector<float> &current, int nb)
{
    std::vector<float> out((size_t)nb * m_a_dim);
    if (mppi_get_new((int)current.size() / m_a_dim, m_a_dim, current.data(), nb, nb ? out.data() : nullptr) != MPPI_OK)
        die("mGetNew");
    return out;
}
