// xform_matrix.cpp -- see xform_matrix.h.
#include "xform_matrix.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace b2r {

Mat33::Mat33()
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            m[i][j] = (i == j) ? 1.0f : 0.0f;
}

Mat33::Mat33(const float* nine) { memcpy(m, nine, sizeof(m)); }

void
Mat33::store(float* nine) const
{
    memcpy(nine, m, sizeof(m));
}

Mat33&
Mat33::translate(float tx, float ty)
{
    for (int j = 0; j < 3; ++j)
        m[2][j] += tx * m[0][j] + ty * m[1][j];
    return *this;
}

Mat33&
Mat33::scale(float sx, float sy)
{
    for (int j = 0; j < 3; ++j) {
        m[0][j] *= sx;
        m[1][j] *= sy;
    }
    return *this;
}

Mat33&
Mat33::operator*=(const Mat33& rhs)
{
    Mat33 out;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < 3; ++k)
                acc += m[i][k] * rhs.m[k][j];
            out.m[i][j] = acc;
        }
    *this = out;
    return *this;
}

Mat33&
Mat33::rotate(float radians)
{
    const float c = cosf(radians), s = sinf(radians);
    Mat33 r;
    r.m[0][0] = c, r.m[0][1] = s;
    r.m[1][0] = -s, r.m[1][1] = c;
    return *this *= r;
}

namespace {
// Divide the leading n x n block of the cofactor matrix by the determinant
// the way Imath guards it: plain division when |det| >= 1, otherwise only if
// no quotient can overflow.
bool
divide_block(Mat33& adj, int n, float det)
{
    if (fabsf(det) >= 1.0f) {
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
                adj.m[i][j] /= det;
        return true;
    }
    const float limit = fabsf(det) / std::numeric_limits<float>::min();
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            if (limit > fabsf(adj.m[i][j]))
                adj.m[i][j] /= det;
            else
                return false;
        }
    return true;
}
}  // namespace

Mat33
Mat33::inverse() const
{
    Mat33 a;
    if (m[0][2] != 0.0f || m[1][2] != 0.0f || m[2][2] != 1.0f) {
        // general (projective) matrix: adjugate / determinant
        a.m[0][0] = m[1][1] * m[2][2] - m[2][1] * m[1][2];
        a.m[0][1] = m[2][1] * m[0][2] - m[0][1] * m[2][2];
        a.m[0][2] = m[0][1] * m[1][2] - m[1][1] * m[0][2];
        a.m[1][0] = m[2][0] * m[1][2] - m[1][0] * m[2][2];
        a.m[1][1] = m[0][0] * m[2][2] - m[2][0] * m[0][2];
        a.m[1][2] = m[1][0] * m[0][2] - m[0][0] * m[1][2];
        a.m[2][0] = m[1][0] * m[2][1] - m[2][0] * m[1][1];
        a.m[2][1] = m[2][0] * m[0][1] - m[0][0] * m[2][1];
        a.m[2][2] = m[0][0] * m[1][1] - m[1][0] * m[0][1];
        const float det = m[0][0] * a.m[0][0] + m[0][1] * a.m[1][0] + m[0][2] * a.m[2][0];
        return divide_block(a, 3, det) ? a : Mat33();
    }
    // affine matrix: invert the 2x2 block, then the translation row
    a.m[0][0] = m[1][1], a.m[0][1] = -m[0][1], a.m[0][2] = 0.0f;
    a.m[1][0] = -m[1][0], a.m[1][1] = m[0][0], a.m[1][2] = 0.0f;
    a.m[2][0] = 0.0f, a.m[2][1] = 0.0f, a.m[2][2] = 1.0f;
    const float det = m[0][0] * m[1][1] - m[1][0] * m[0][1];
    if (!divide_block(a, 2, det))
        return Mat33();
    a.m[2][0] = -m[2][0] * a.m[0][0] - m[2][1] * a.m[1][0];
    a.m[2][1] = -m[2][0] * a.m[0][1] - m[2][1] * a.m[1][1];
    return a;
}

void
Mat33::transform_point(float x, float y, float* ox, float* oy) const
{
    const float a = x * m[0][0] + y * m[1][0] + m[2][0];
    const float b = x * m[0][1] + y * m[1][1] + m[2][1];
    const float w = x * m[0][2] + y * m[1][2] + m[2][2];
    *ox           = a / w;
    *oy           = b / w;
}

Mat33
rotation_about(float radians, float cx, float cy)
{
    Mat33 M;
    M.translate(-cx, -cy);
    M.rotate(radians);
    Mat33 back;
    back.translate(cx, cy);
    M *= back;
    return M;
}

Mat33
fromto_matrix(float from_x, float from_y, float from_w, float from_h, float to_x, float to_y,
              float to_w, float to_h)
{
    Mat33 M;
    M.translate(to_x, to_y);
    M.scale(to_w / from_w, to_h / from_h);
    M.translate(-from_x, -from_y);
    return M;
}

void
transform_bounds(const Mat33& M, int xbegin, int xend, int ybegin, int yend, int out[4])
{
    const float cx[2] = { xbegin + 0.5f, xend - 0.5f };
    const float cy[2] = { ybegin + 0.5f, yend - 0.5f };
    float lo_x = 0, hi_x = 0, lo_y = 0, hi_y = 0;
    bool first = true;
    for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 2; ++i) {
            float px, py;
            M.transform_point(cx[i], cy[j], &px, &py);
            if (first) {
                lo_x = hi_x = px, lo_y = hi_y = py;
                first = false;
            } else {
                lo_x = std::min(lo_x, px), hi_x = std::max(hi_x, px);
                lo_y = std::min(lo_y, py), hi_y = std::max(hi_y, py);
            }
        }
    out[0] = int(floorf(lo_x));
    out[1] = int(floorf(hi_x)) + 1;
    out[2] = int(floorf(lo_y));
    out[3] = int(floorf(hi_y)) + 1;
}

}  // namespace b2r
