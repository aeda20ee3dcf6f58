// Test-infrastructure shim; see snn.h in this directory.
#include "snn.h"

#include <algorithm>
#include <cmath>
#include <numeric>

SnnModel::SnnModel(double* data, int r, int c) : rows(r), cols(c) {
    mu.assign(cols, 0.0);
    for (int j = 0; j < cols; ++j) {
        double s = 0.0;
        for (int i = 0; i < rows; ++i) s += data[i + static_cast<std::size_t>(rows) * j];
        mu[j] = s / static_cast<double>(rows);
    }
    std::vector<double> X(static_cast<std::size_t>(rows) * cols);
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) X[static_cast<std::size_t>(i) * cols + j] = data[i + static_cast<std::size_t>(rows) * j] - mu[j];

    axis.assign(cols, 0.0);
    axis[0] = 1.0;
    if (cols > 1 && rows > 1) {
        std::vector<double> G(static_cast<std::size_t>(cols) * cols, 0.0);
        for (int i = 0; i < rows; ++i)
            for (int a = 0; a < cols; ++a)
                for (int b = 0; b < cols; ++b) G[a * cols + b] += X[static_cast<std::size_t>(i) * cols + a] * X[static_cast<std::size_t>(i) * cols + b];
        std::vector<double> v(cols, 1.0 / std::sqrt(static_cast<double>(cols))), w(cols);
        for (int it = 0; it < 64; ++it) {
            double nrm = 0.0;
            for (int a = 0; a < cols; ++a) {
                double s = 0.0;
                for (int b = 0; b < cols; ++b) s += G[a * cols + b] * v[b];
                w[a] = s;
                nrm += s * s;
            }
            nrm = std::sqrt(nrm);
            if (!(nrm > 0.0)) break;
            for (int a = 0; a < cols; ++a) v[a] = w[a] / nrm;
        }
        double nrm = 0.0;
        for (double e : v) nrm += e * e;
        if (nrm > 0.0) {
            nrm = std::sqrt(nrm);
            for (int a = 0; a < cols; ++a) axis[a] = v[a] / nrm;
        }
    }

    std::vector<double> proj(rows);
    for (int i = 0; i < rows; ++i) {
        double s = 0.0;
        for (int j = 0; j < cols; ++j) s += X[static_cast<std::size_t>(i) * cols + j] * axis[j];
        proj[i] = s;
    }
    order.resize(rows);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return proj[a] < proj[b]; });

    key.resize(rows);
    sq.resize(rows);
    centred.resize(static_cast<std::size_t>(rows) * cols);
    for (int i = 0; i < rows; ++i) {
        key[i] = proj[order[i]];
        double s = 0.0;
        for (int j = 0; j < cols; ++j) {
            const double e = X[static_cast<std::size_t>(order[i]) * cols + j];
            centred[static_cast<std::size_t>(i) * cols + j] = e;
            s += e * e;
        }
        sq[i] = s;
    }
}

std::pair<std::size_t, std::size_t> SnnModel::window(double proj, double radius) const {
    // widen by a few ulps so rounding in the projection can never drop an in-radius point
    const double slack = 1e-9 * (std::abs(proj) + radius + 1.0);
    const auto lo = std::lower_bound(key.begin(), key.end(), proj - radius - slack) - key.begin();
    const auto hi = std::upper_bound(key.begin(), key.end(), proj + radius + slack) - key.begin();
    return {static_cast<std::size_t>(lo), static_cast<std::size_t>(hi)};
}
