// ModelBase — point-mass LTI model x' = A x + (B/m) u on the GPU.
// Same class name and constructor as /root/reference/include/model_base.hpp:10-134; the
// TensorFlow graph builders (mBuild*Graph taking Scope/Input) become batched calls on plain
// buffers that run the CUDA stage kernels through the C-ABI (include/mppi_b200.h).
#ifndef MPPI_B200_MODEL_BASE_HPP
#define MPPI_B200_MODEL_BASE_HPP

#include <vector>

class ModelBase {
public:
    ModelBase();                                                       // model_base.hpp:43, defaults of src/model_base.cpp:12: m=1, dt=.01, s=2, a=1
    ModelBase(const float mass, const float dt, const int s_dim, const int a_dim);   // :58-61
    ~ModelBase();

    // mBuildModelStepGraph (:85-87): state [k|1][s], action [k][a] -> next state [k][s].
    // A leading state dimension of 1 broadcasts (test/test_model.cpp:216-255).
    std::vector<float> predict(const std::vector<float> &state, const std::vector<float> &action) const;
    // mBuildFreeStepGraph (:101-102): A x, state [k][s] -> [k][s]
    std::vector<float> freeStep(const std::vector<float> &state) const;
    // mBuildActionStepGraph (:116-117): (B/m) u, action [k][a] -> [k][s]
    std::vector<float> actionStep(const std::vector<float> &action) const;
    // A = blockDiag([[1,dt],[0,1]], s/2), B = blockDiag([[dt^2/2],[dt]]/m, a) as row-major matrices
    std::vector<float> A() const;
    std::vector<float> B() const;
    void train();                                                      // :130, empty in the reference too

    float mass() const { return m_m; }
    float dt() const { return m_dt; }
    int sDim() const { return m_s_dim; }
    int aDim() const { return m_a_dim; }
    void setDevice(int device) { m_device = device; }

private:
    float m_dt, m_m;
    int m_s_dim, m_a_dim;
    int m_device = -1;
};

#endif
