// shim_arma/mlpack/core.hpp — mlpack::data::Load for a headerless numeric CSV (TEST INFRASTRUCTURE; see shim_arma/armadillo)
#pragma once
#include <armadillo>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
namespace mlpack { namespace data {
// transpose == false: the matrix as the file has it (rows = records), which is how BIC_OLS.cpp:48 calls it
inline bool Load(const std::string &filename, arma::mat &m, bool fatal = false, bool transpose = true) {
    std::ifstream in(filename);
    if (!in.good()) { if (fatal) throw std::runtime_error("Cannot open file '" + filename + "'"); return false; }
    std::vector<std::vector<double>> rows;
    std::string line;
    while (std::getline(in, line)) {
        if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;
        std::vector<double> r;
        std::stringstream ss(line);
        std::string tok;
        while (std::getline(ss, tok, ',')) r.push_back(strtod(tok.c_str(), nullptr));
        rows.push_back(r);
    }
    const arma::uword nr = rows.size(), nc = nr ? rows[0].size() : 0;
    arma::mat a(nr, nc);
    for (arma::uword i = 0; i < nr; i++)
        for (arma::uword j = 0; j < nc && j < rows[i].size(); j++) a(i, j) = rows[i][j];
    m = transpose ? arma::trans(a) : a;
    return true;
}
} }
