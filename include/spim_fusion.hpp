// spim_fusion.hpp -- header-only C++ host layer above include/spim_fusion.h, mirroring the reference's Java classes of
// the fusion pre-step with the same names and argument meaning (paths under /root/reference/src/main/java/):
//
//   spim/process/fusion/deconvolution/ProcessForDeconvolution.java  (fuseStacksAndGetPSFs, WeightType, OSEM factor)
//   spim/process/fusion/deconvolution/ExtractPSF.java               (extractNextImg, transformPSF)
//   spim/process/fusion/weights/Blending.java
//   net.imglib2.realtransform.AffineTransform3D                     (set / inverse / getRowPackedCopy only)
//
// Unlike the Java original, the transformed views stay on the GPU inside the session that will deconvolve them
// (getTransformedImg / getTransformedWeight download a copy on request); deconvolve() then runs the iterations on the
// resident data.  Coordinates, offsets and affines are in (x, y, z) order like the Java code; Image is [z][y][x].
#pragma once
#include "spim_fusion.h"
#include "spim_mvdecon.hpp"

#include <algorithm>
#include <cmath>

namespace spim_b200 {

/// ProcessForDeconvolution.java:81 (ordinal order)
enum class WeightType : int { WEIGHTS_ONLY = 0, NO_WEIGHTS = 1, VIRTUAL_WEIGHTS = 2, PRECOMPUTED_WEIGHTS = 3, LOAD_WEIGHTS = 4 };

class AffineTransform3D {
public:
    AffineTransform3D() : m_{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}} {}
    explicit AffineTransform3D(const std::array<double, 12>& rowPacked) : m_(rowPacked) {}
    void set(const std::array<double, 12>& rowPacked) { m_ = rowPacked; }
    const std::array<double, 12>& getRowPackedCopy() const { return m_; }
    /// AffineTransform3D.invert(): adjugate / determinant, translation = -(R^-1 t)
    AffineTransform3D inverse() const {
        const double m00 = m_[0], m01 = m_[1], m02 = m_[2], m03 = m_[3], m10 = m_[4], m11 = m_[5], m12 = m_[6], m13 = m_[7],
                     m20 = m_[8], m21 = m_[9], m22 = m_[10], m23 = m_[11];
        const double det = m00 * m11 * m22 + m10 * m21 * m02 + m20 * m01 * m12 - m02 * m11 * m20 - m12 * m21 * m00 - m22 * m01 * m10;
        if (det == 0) throw std::runtime_error("Matrix is singular.");
        const double idet = 1.0 / det;
        std::array<double, 12> i;
        i[0] = (m11 * m22 - m12 * m21) * idet; i[1] = (m02 * m21 - m01 * m22) * idet; i[2] = (m01 * m12 - m02 * m11) * idet;
        i[4] = (m12 * m20 - m10 * m22) * idet; i[5] = (m00 * m22 - m02 * m20) * idet; i[6] = (m02 * m10 - m00 * m12) * idet;
        i[8] = (m10 * m21 - m11 * m20) * idet; i[9] = (m01 * m20 - m00 * m21) * idet; i[10] = (m00 * m11 - m01 * m10) * idet;
        i[3] = -i[0] * m03 - i[1] * m13 - i[2] * m23;
        i[7] = -i[4] * m03 - i[5] * m13 - i[6] * m23;
        i[11] = -i[8] * m03 - i[9] * m13 - i[10] * m23;
        return AffineTransform3D(i);
    }
private:
    std::array<double, 12> m_;
};

/// Blending( interval, border, blending ) -- weights/Blending.java:44-50; the interval is the raw stack
struct Blending {
    std::array<float, 3> border{{0, 0, 0}}, blending{{0, 0, 0}};
};

/// ExtractPSF.java: PSFs per view, extracted from the stack that is loaded in a session
class ExtractPSF {
public:
    /// transformPSF (ExtractPSF.java:325-367)
    static Image transformPSF(const Image& psf, const AffineTransform3D& model, int device = 0) {
        const int d[3] = {psf.dims[2], psf.dims[1], psf.dims[0]};
        int od[3]; double off[3];
        check(mvd_transform_psf_size(d, model.getRowPackedCopy().data(), od, off), "mvd_transform_psf_size");
        Image out(od[2], od[1], od[0]);
        check(mvd_transform_psf(psf.data.data(), d, model.getRowPackedCopy().data(), model.inverse().getRowPackedCopy().data(),
                                out.data.data(), od, device), "mvd_transform_psf");
        return out;
    }
    /// extractNextImg (ExtractPSF.java:277-296) on the stack loaded in `s`; locations are (x, y, z) triples, psfSize (x, y, z)
    void extractNextImg(mvd_session* s, int viewId, const AffineTransform3D& model, const std::vector<std::array<double, 3>>& locations,
                        const std::array<int, 3>& psfSize, int device = 0) {
        Image original(psfSize[0], psfSize[1], psfSize[2]);
        const int sz[3] = {psfSize[2], psfSize[1], psfSize[0]};
        check(mvd_extract_psf(s, (int)locations.size(), locations.empty() ? nullptr : locations[0].data(), sz, 1, original.data.data()),
              "mvd_extract_psf");
        viewIds_.push_back(viewId);
        psfs_.push_back(transformPSF(original, model, device));
        originals_.push_back(std::move(original));
    }
    const Image& getTransformedPSF(int viewId) const {
        for (size_t i = 0; i < viewIds_.size(); ++i) if (viewIds_[i] == viewId) return psfs_[i];
        throw std::runtime_error("Cannot find PSF for view " + std::to_string(viewId));
    }
    const Image& getInputCalibrationPSF(int viewId) const {
        for (size_t i = 0; i < viewIds_.size(); ++i) if (viewIds_[i] == viewId) return originals_[i];
        throw std::runtime_error("Cannot find PSF for view " + std::to_string(viewId));
    }
    const std::vector<int>& getViewIdsForPSFs() const { return viewIds_; }
private:
    std::vector<int> viewIds_;
    std::vector<Image> psfs_, originals_;
};

/// ProcessForDeconvolution.java:96-366 + the deconvolution it prepares (EfficientBayesianBased.java:254-285)
class ProcessForDeconvolution {
public:
    /// bbDims / bbMin: the bounding box (x, y, z); numThreads: Threads.numThreads() of the reference (portion count of the
    /// overlap statistics)
    ProcessForDeconvolution(std::array<int, 3> bbDims, std::array<long long, 3> bbMin, std::array<int, 3> blendingBorder,
                            std::array<int, 3> blendingRange, int numViews, PSFTYPE iterationType, double lambda,
                            int osemIndex = 0, double osemspeedup = 1.0, int numThreads = 1, int device = 0)
        : bbDims_(bbDims), bbMin_(bbMin), border_(blendingBorder), range_(blendingRange), numViews_(numViews),
          osemIndex_(osemIndex), osemspeedup_(osemspeedup), numThreads_(numThreads), device_(device) {
        mvd_params p;
        mvd_params_default(&p);
        p.dims[0] = bbDims[2]; p.dims[1] = bbDims[1]; p.dims[2] = bbDims[0];
        p.num_views = numViews;
        p.iteration_type = (int)iterationType;
        p.generation = 2;
        p.lambda = lambda;
        p.osem_speedup = osemspeedup;
        p.osem_index = osemIndex;
        p.device = device;
        check(mvd_session_create(&p, &s_), "mvd_session_create");
    }
    ~ProcessForDeconvolution() { if (s_) mvd_session_destroy(s_); }
    ProcessForDeconvolution(const ProcessForDeconvolution&) = delete;
    ProcessForDeconvolution& operator=(const ProcessForDeconvolution&) = delete;

    /// stacks: raw image stacks (min-max normalised on the device like ProcessFusion.getImage(..., true));
    /// psfs: transformed PSFs, or empty when bead locations are given for extraction
    bool fuseStacksAndGetPSFs(const std::vector<Image>& stacks, const std::vector<AffineTransform3D>& transforms, WeightType weightType,
                              const std::vector<Image>& psfs, const std::vector<std::vector<std::array<double, 3>>>& beads = {},
                              std::array<int, 3> psfSize = {{0, 0, 0}}) {
        if (stacks.empty() || (int)stacks.size() != numViews_ || transforms.size() != stacks.size()) return false;
        if (weightType == WeightType::LOAD_WEIGHTS) throw std::runtime_error("LOAD_WEIGHTS not implemented yet.");
        const bool extract = !beads.empty();
        if (!extract && psfs.size() != stacks.size() && weightType != WeightType::WEIGHTS_ONLY) return false;
        for (int v = 0; v < numViews_; ++v) {
            const Image& st = stacks[v];
            const int sd[3] = {st.dims[2], st.dims[1], st.dims[0]};
            check(mvd_load_stack(s_, st.data.data(), sd, 1), "mvd_load_stack");
            mvd_transform t;
            std::fill((char*)&t, (char*)&t + sizeof(t), 0);
            t.struct_size = (int)sizeof(t);
            const auto inv = transforms[v].inverse().getRowPackedCopy();
            for (int i = 0; i < 12; ++i) t.inverse[i] = inv[i];
            for (int d = 0; d < 3; ++d) { t.offset[d] = bbMin_[d]; t.border[d] = (float)border_[d]; t.range[d] = (float)range_[d]; }
            t.want_image = weightType != WeightType::WEIGHTS_ONLY;
            t.want_weight = weightType != WeightType::NO_WEIGHTS;
            check(mvd_transform_view(s_, v, &t), "mvd_transform_view");
            if (extract) ePSF_.extractNextImg(s_, v, transforms[v], beads[v], psfSize, device_);
        }
        check(mvd_load_stack(s_, nullptr, nullptr, 0), "mvd_load_stack");
        for (int v = 0; v < numViews_; ++v) {
            if (!extract && psfs.empty()) break;
            const Image& k = extract ? ePSF_.getTransformedPSF(v) : psfs[v];
            const int kd[3] = {k.dims[2], k.dims[1], k.dims[0]};
            check(mvd_set_psf(s_, v, k.data.data(), kd), "mvd_set_psf");
        }
        if (weightType != WeightType::NO_WEIGHTS) {
            int mn = 0; double avg = 0;
            check(mvd_normalize_weights(s_, weightType == WeightType::VIRTUAL_WEIGHTS ? MVD_WEIGHTS_VIRTUAL : MVD_WEIGHTS_PRECOMPUTED,
                                        numThreads_ * 2, &mn, &avg), "mvd_normalize_weights");
            minOverlappingViews_ = std::max(1, mn);                 // ProcessForDeconvolution.java:349-350
            avgOverlappingViews_ = std::max(1.0, avg);
        }
        if (osemIndex_ == 1) osemspeedup_ = minOverlappingViews_;
        else if (osemIndex_ == 2) osemspeedup_ = avgOverlappingViews_;
        return true;
    }

    ExtractPSF& getExtractPSF() { return ePSF_; }
    int getMinOverlappingViews() const { return minOverlappingViews_; }
    double getAvgOverlappingViews() const { return avgOverlappingViews_; }
    double getOSEMspeedup() const { return osemspeedup_; }
    Image getTransformedImg(int v) { return fetch(v, 0); }
    Image getTransformedWeight(int v) { return fetch(v, 1); }

    /// new MVDeconvolution( deconvolutionData, iterationType, numIterations, lambda, ... ).getPsi() on the resident views
    Image deconvolve(int numIterations) {
        check(mvd_init(s_), "mvd_init");
        check(mvd_run(s_, numIterations, nullptr, nullptr), "mvd_run");
        check(mvd_finish(s_), "mvd_finish");
        Image out(bbDims_[0], bbDims_[1], bbDims_[2]);
        check(mvd_get_psi(s_, out.data.data()), "mvd_get_psi");
        return out;
    }
    mvd_session* session() { return s_; }

private:
    Image fetch(int v, int which) {
        Image out(bbDims_[0], bbDims_[1], bbDims_[2]);
        check(mvd_get_view(s_, v, which, out.data.data()), "mvd_get_view");
        return out;
    }
    std::array<int, 3> bbDims_;
    std::array<long long, 3> bbMin_;
    std::array<int, 3> border_, range_;
    int numViews_, osemIndex_;
    double osemspeedup_;
    int numThreads_, device_;
    mvd_session* s_ = nullptr;
    ExtractPSF ePSF_;
    int minOverlappingViews_ = 0;
    double avgOverlappingViews_ = 0.0;
};

}  // namespace spim_b200
