// Test-infrastructure shim (NOT product code, NOT reference code).
//
// Eigen-free stand-in for the vendored third-party SNN model (src/SNN/include/snn.h:30-85,
// src/SNN/src/snn.cpp:97-160), which needs Eigen::BDCSVD.  It keeps the public surface the
// reference's SNNQueries wrapper uses (SNNQueries.cpp:20,35) and SNN's published algorithm:
// centre the data, project on a unit axis, sort by the projection, and answer a radius
// query by scanning the projection window [q-r, q+r] with exact squared distances
// evaluated as ||x||^2 + ||q||^2 - 2 x.q on the centred data.  The axis is the dominant
// eigenvector of X^T X by power iteration instead of a BDCSVD row; any unit axis gives
// the same result set ("all points within r"), only the output order may differ.
#pragma once
#include <cstddef>
#include <utility>
#include <vector>

class SnnModel {
   public:
    class Vector {
       public:
        std::size_t size() const { return v_.size(); }
        void resize(std::size_t n) { v_.resize(n); }
        double& operator[](std::size_t i) { return v_[i]; }
        double operator[](std::size_t i) const { return v_[i]; }
       private:
        std::vector<double> v_;
    };

    SnnModel() = default;
    SnnModel(double* columnMajor, int r, int c);

    SnnModel(const SnnModel&) = delete;
    SnnModel& operator=(const SnnModel&) = delete;
    SnnModel(SnnModel&&) = default;
    SnnModel& operator=(SnnModel&&) = default;

    template <typename InputVectorT, typename ResultT, typename ResultMappingFn>
    void radius_single_query(const InputVectorT& query, double radius, std::vector<ResultT>& out,
                             ResultMappingFn mapping, Vector& qbuf, Vector& dbuf) const {
        if (qbuf.size() < static_cast<std::size_t>(cols)) qbuf.resize(cols);
        (void)dbuf;
        double qq = 0.0, proj = 0.0;
        for (int j = 0; j < cols; ++j) {
            qbuf[j] = query[j] - mu[j];
            qq += qbuf[j] * qbuf[j];
            proj += qbuf[j] * axis[j];
        }
        const auto [lo, hi] = window(proj, radius);
        const double r2 = radius * radius;
        for (std::size_t i = lo; i < hi; ++i) {
            const double* row = &centred[i * cols];
            double dot = 0.0;
            for (int j = 0; j < cols; ++j) dot += row[j] * qbuf[j];
            if (sq[i] + qq - 2.0 * dot <= r2) out.push_back(mapping(order[i]));
        }
    }

   private:
    std::pair<std::size_t, std::size_t> window(double proj, double radius) const;

    int rows = 0, cols = 0;
    std::vector<double> mu, axis, key, centred, sq;  // centred: rows x cols row-major, sorted by key
    std::vector<int> order;
};
