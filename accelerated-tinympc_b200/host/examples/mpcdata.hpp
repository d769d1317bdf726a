// Reader for the repo's .mpcdata problem files (accelerated-tinympc_b200/problem_data): "scalar name value" and
// "matrix name rows cols" followed by one line of column-major values.
#pragma once
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

struct MpcData {
    std::map<std::string, double> scalars;
    struct M { int rows, cols; std::vector<double> v; };   // column-major
    std::map<std::string, M> mats;

    explicit MpcData(const std::string &path)
    {
        std::ifstream f(path);
        if (!f) throw std::runtime_error("cannot open " + path);
        std::string line;
        while (std::getline(f, line)) {
            if (line.empty() || line[0] == '#') continue;
            std::istringstream is(line);
            std::string kind, name;
            is >> kind >> name;
            if (kind == "scalar") {
                double v; is >> v; scalars[name] = v;
            } else if (kind == "matrix") {
                M m; is >> m.rows >> m.cols;
                std::string vals;
                std::getline(f, vals);
                std::istringstream vs(vals);
                double x;
                while (vs >> x) m.v.push_back(x);
                if ((int)m.v.size() != m.rows * m.cols) throw std::runtime_error("bad matrix " + name + " in " + path);
                mats[name] = m;
            }
        }
    }
    template <class T> std::vector<T> cast(const std::string &name) const
    {
        const M &m = mats.at(name);
        return std::vector<T>(m.v.begin(), m.v.end());
    }
};
