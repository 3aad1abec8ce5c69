// DataBase: host-side transition log (behaviour of /root/reference/src/data_base.cpp:14-71).
#include "data_base.hpp"

#include <fstream>
#include <iostream>

DataBase::DataBase() {}
DataBase::~DataBase() {}

void DataBase::addX(std::vector<float> x) { state_input.push_back(std::move(x)); }
void DataBase::addU(std::vector<float> u) { action_input.push_back(std::move(u)); }
void DataBase::addNext(std::vector<float> x_next) { output.push_back(std::move(x_next)); }

void DataBase::addEl(std::vector<float> x, std::vector<float> u, std::vector<float> x_next)
{
    addX(std::move(x));
    addU(std::move(u));
    addNext(std::move(x_next));
}

std::string DataBase::tensor2CSV(const std::vector<float> &t)
{
    std::string line;
    for (float v : t) line += std::to_string(v) + ",";
    return line;
}

std::string DataBase::csvHeader(const std::vector<float> &t, std::string prefix)
{
    std::string line;
    for (size_t i = 0; i < t.size(); i++) line += prefix + std::to_string(i) + ",";
    return line;
}

void DataBase::toCSV(std::string filename)
{
    if (state_input.size() != action_input.size() && state_input.size() != output.size())
        std::cerr << "The vector size don't match" << std::endl;
    std::ofstream out(filename);
    if (state_input.empty() || action_input.empty() || output.empty()) return;
    out << csvHeader(state_input[0], "x") << csvHeader(action_input[0], "u") << csvHeader(output[0], "x_next")
        << std::endl;
    const size_t n = std::min(state_input.size(), std::min(action_input.size(), output.size()));
    for (size_t i = 0; i < n; i++)
        out << tensor2CSV(state_input[i]) << tensor2CSV(action_input[i]) << tensor2CSV(output[i]) << std::endl;
}
