// DataBase — transition log (x, u, x_next) with CSV export.
// Same public interface as /root/reference/include/data_base.hpp:8-50 with tensorflow::Tensor
// replaced by std::vector<float>; file format identical to src/data_base.cpp:52-71
// ("x0,x1,..,u0,..,x_next0,..," header, std::to_string values, trailing commas).
#ifndef MPPI_B200_DATA_BASE_HPP
#define MPPI_B200_DATA_BASE_HPP

#include <string>
#include <vector>

class DataBase {
public:
    DataBase();
    ~DataBase();

    void addEl(std::vector<float> x, std::vector<float> u, std::vector<float> x_next);
    void addX(std::vector<float> x);
    void addU(std::vector<float> u);
    void addNext(std::vector<float> x_next);
    void toCSV(std::string filename);
    std::string csvHeader(const std::vector<float> &t, std::string prefix);
    std::string tensor2CSV(const std::vector<float> &t);
    size_t size() const { return state_input.size(); }

private:
    std::vector<std::vector<float>> state_input;
    std::vector<std::vector<float>> action_input;
    std::vector<std::vector<float>> output;
};

#endif
