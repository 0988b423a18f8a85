// xform_matrix.h -- the 3x3 float matrix algebra ImageBufAlgo::warp / rotate /
// fit(exact) and oiiotool --resize:from/to need on the host.
//
// The reference uses Imath::M33f (row-vector convention: p' = p * M,
// translation in row 2).  Imath is not part of the reference tree; only the
// handful of operations the path uses are provided, each computed in the
// same order of float operations as Imath 3.1's Matrix33<float> so the
// matrices -- and through them the footprints -- agree with the reference
// (compile with -ffp-contract=off).
#pragma once

namespace b2r {

struct Mat33 {
    float m[3][3];

    Mat33();  // identity
    explicit Mat33(const float* nine);
    void store(float* nine) const;

    Mat33& translate(float tx, float ty);  // Matrix33::translate(V2)
    Mat33& scale(float sx, float sy);      // Matrix33::scale(V2)
    Mat33& rotate(float radians);          // Matrix33::rotate(r)
    Mat33& operator*=(const Mat33& rhs);
    Mat33 inverse() const;  // identity when singular (non-throwing form)
    // multVecMatrix: (x, y, 1) * M with the homogeneous divide
    void transform_point(float x, float y, float* ox, float* oy) const;
};

// ImageBufAlgo::rotate's matrix (imagebufalgo_xform.cpp:1733-1737)
Mat33
rotation_about(float radians, float cx, float cy);
// oiiotool --resize:from=:to= (oiiotool.cpp:4704-4707)
Mat33
fromto_matrix(float from_x, float from_y, float from_w, float from_h, float to_x, float to_y,
              float to_w, float to_h);
// transform(M, roi) (imagebufalgo_xform.cpp:231-253): integer bounding box of
// the transformed pixel-centre corners
void
transform_bounds(const Mat33& M, int xbegin, int xend, int ybegin, int yend, int out[4]);

}  // namespace b2r
