// spim_mvdecon.hpp -- header-only C++ host layer above the C-ABI (include/spim_mvdecon.h,
// include/spim_fftconv.h), mirroring the reference's Java classes for the deconvolution path with the
// same names, constructor arguments, argument meaning and error behaviour:
//
//   gen-1  mpicbg/spim/postprocessing/deconvolution2/{LRFFT,LRInput,BayesMVDeconvolution,Deconvolver}.java
//   gen-2  spim/process/fusion/deconvolution/{MVDeconFFT,MVDeconInput,MVDeconvolution}.java
//
// The reference is Java and there is no JVM in the build image, so this is what a host application
// links against (the JNA stubs for the Java side are in java/).  Volumes are dense float arrays in
// [z][y][x] order with dims given as (x, y, z) like the Java classes (ImgLib dimension order); the
// layer reverses them to (z, y, x) for the native calls exactly as getCUDACoordinates does
// (MVDeconFFTThreads.java:157-165).  Everything computes on the GPU through the shared library; a
// failing native call throws std::runtime_error with mvd_last_error().
#pragma once
#include "spim_mvdecon.h"
#include "spim_fftconv.h"

#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace spim_b200 {

/// LRFFT.java:54 / MVDeconFFT.java:49 -- ordinal order matters
enum class PSFTYPE : int { OPTIMIZATION_II = 0, OPTIMIZATION_I = 1, EFFICIENT_BAYESIAN = 2, INDEPENDENT = 3 };

constexpr float minValue = 0.0001f;   // LRInput.java:30, MVDeconvolution.java:70

/// dense fp32 volume, dims (x, y, z), data [z][y][x]
struct Image {
    std::array<int, 3> dims{{0, 0, 0}};
    std::vector<float> data;
    Image() = default;
    Image(int x, int y, int z, float fill = 0.f) : dims{{x, y, z}}, data((size_t)x * y * z, fill) {}
    size_t size() const { return (size_t)dims[0] * dims[1] * dims[2]; }
    bool empty() const { return data.empty(); }
};

inline void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + mvd_last_error());
}

/// common part of LRFFT (LRFFT.java:58-204) and MVDeconFFT (MVDeconFFT.java:53-172)
class ViewFFT {
public:
    ViewFFT(Image image, Image weight, Image kernel, std::vector<int> deviceList, bool useBlocks,
            std::array<int, 3> blockSize)
        : image_(std::move(image)), weight_(std::move(weight)), kernel1_(std::move(kernel)),
          deviceList_(std::move(deviceList)), useBlocks_(useBlocks), blockSize_(blockSize) {
        if (deviceList_.empty()) deviceList_.push_back(0);
        for (int d : deviceList_)
            if (d < 0) throw std::invalid_argument("device id -1 (CPU) is not supported: there is no CPU path");
        if (!weight_.empty() && weight_.dims != image_.dims) throw std::invalid_argument("weight dims != image dims");
    }
    virtual ~ViewFFT() = default;

    void setNumViews(int n) { numViews_ = n; }
    const Image& getImage() const { return image_; }
    const Image& getWeight() const { return weight_; }     // empty = constant 1 (LRFFT.java:201-204)
    const Image& getKernel1() const { return kernel1_; }
    const Image& getKernel2() const { return kernel2_; }
    void setImage(Image i) { image_ = std::move(i); iteration_ = -1; }
    void setWeight(Image w) { weight_ = std::move(w); }
    void setCurrentIteration(int i) { iteration_ = i; }
    int getCurrentIteration() const { return iteration_; }
    int device0() const { return deviceList_[0]; }

    /// psi (*) kernel1 with mirror extension through the legacy entry, one single block
    /// (LRFFT.java:423-526 / MVDeconFFT.java:384-470, CUDA one-device path)
    Image convolve1(const Image& psi) const { return conv(psi, kernel1_, MVD_EXT_MIRROR_SINGLE, 0.f); }
    /// ratio (*) kernel2; gen-2 extends with 1.0 (MVDeconFFT.java:515)
    Image convolve2(const Image& ratio) const {
        if (kernel2_.empty()) throw std::runtime_error("kernel2 not initialised: call init on the input first");
        return generation() == 2 ? conv(ratio, kernel2_, MVD_EXT_CONSTANT, 1.f) : conv(ratio, kernel2_, MVD_EXT_MIRROR_SINGLE, 0.f);
    }

protected:
    virtual int generation() const = 0;
    Image conv(const Image& in, const Image& k, int ext, float value) const {
        Image out(in.dims[0], in.dims[1], in.dims[2]);
        const int id[3] = {in.dims[2], in.dims[1], in.dims[0]};
        const int kd[3] = {k.dims[2], k.dims[1], k.dims[0]};
        check(mvd_convolve(in.data.data(), id, k.data.data(), kd, ext, value, out.data.data(), device0()), "mvd_convolve");
        return out;
    }
    Image image_, weight_, kernel1_, kernel2_;
    std::vector<int> deviceList_;
    bool useBlocks_;
    std::array<int, 3> blockSize_;
    int numViews_ = 0;
    int iteration_ = -1;
    template <class V> friend class Input;
    template <class V, int G> friend class Deconvolution;
};

/// LRFFT( image, weight, kernel, deviceList, useBlocks, blockSize )  -- LRFFT.java:131-199
class LRFFT : public ViewFFT {
public:
    using ViewFFT::ViewFFT;
protected:
    int generation() const override { return 1; }
};

/// MVDeconFFT( image, weight, kernel, blockFactory, deviceList, useBlocks, blockSize, saveMemory ) -- MVDeconFFT.java:79-85
/// (the ImgLib factory argument has no meaning here and is dropped)
class MVDeconFFT : public ViewFFT {
public:
    MVDeconFFT(Image image, Image weight, Image kernel, std::vector<int> deviceList = {0}, bool useBlocks = false,
               std::array<int, 3> blockSize = {{0, 0, 0}}, bool saveMemory = false)
        : ViewFFT(std::move(image), std::move(weight), std::move(kernel), std::move(deviceList), useBlocks, blockSize),
          saveMemory_(saveMemory) {}
protected:
    int generation() const override { return 2; }
    bool saveMemory_;
};

/// LRInput (LRInput.java:28-76) / MVDeconInput (MVDeconInput.java:31-80)
template <class V>
class Input {
public:
    void add(std::shared_ptr<V> view) {            // re-broadcasts numViews to every view (LRInput.java:33-39)
        views_.push_back(std::move(view));
        for (auto& v : views_) v->setNumViews(getNumViews());
    }
    std::vector<std::shared_ptr<V>>& getViews() { return views_; }
    int getNumViews() const { return (int)views_.size(); }
private:
    std::vector<std::shared_ptr<V>> views_;
};
using LRInput = Input<LRFFT>;
using MVDeconInput = Input<MVDeconFFT>;

/// Deconvolver.java:27-35
class Deconvolver {
public:
    virtual ~Deconvolver() = default;
    virtual std::string getName() const = 0;
    virtual double getAvg() const = 0;
    virtual Image getPsi() = 0;
    virtual void runIteration() = 0;
};

struct ViewStat { int iteration, view; double sumChange, maxChange; };

/// BayesMVDeconvolution (gen-1, BayesMVDeconvolution.java:79-178) and MVDeconvolution (gen-2,
/// MVDeconvolution.java:94-211): like the Java constructors, this one runs every iteration.
template <class V, int GEN>
class Deconvolution : public Deconvolver {
public:
    Deconvolution(Input<V>& views, PSFTYPE iterationType, int numIterations, double lambda, double osemspeedup = 1.0,
                  int osemspeedupindex = 0, std::string name = "deconvolved", bool exactTikhonov = false)
        : views_(views), name_(std::move(name)) {
        auto& data = views.getViews();
        if (data.empty()) throw std::invalid_argument("no views");
        const Image& first = data[0]->getImage();
        mvd_params p;
        mvd_params_default(&p);
        p.dims[0] = first.dims[2]; p.dims[1] = first.dims[1]; p.dims[2] = first.dims[0];
        p.num_views = (int)data.size();
        p.iteration_type = (int)iterationType;
        p.generation = GEN;
        p.lambda = lambda;
        // gen-2 clamps weights upstream; its constructor's osem arguments are unused (MVDeconvolution.java)
        p.osem_speedup = GEN == 1 ? osemspeedup : 1.0;
        p.osem_index = GEN == 1 ? osemspeedupindex : 0;
        p.device = data[0]->device0();
        p.exact_tikhonov = exactTikhonov ? 1 : 0;
        check(mvd_session_create(&p, &s_), "mvd_session_create");
        try {
            for (int v = 0; v < (int)data.size(); ++v) {
                const V& view = *data[v];
                if (view.getImage().dims != first.dims) throw std::invalid_argument("all views must have the same dims");
                const Image& k = view.getKernel1();
                const int kd[3] = {k.dims[2], k.dims[1], k.dims[0]};
                check(mvd_set_view(s_, v, view.getImage().data.data(),
                                   view.getWeight().empty() ? nullptr : view.getWeight().data.data(), k.data.data(), kd),
                      "mvd_set_view");
            }
            check(mvd_init(s_), "mvd_init");              // views.init( iterationType ), psi = avg, OSEM clamp
            for (int v = 0; v < (int)data.size(); ++v) {   // expose the normalised kernel1 and kernel2 on the views
                V& view = *data[v];
                Image k1 = view.getKernel1(), k2 = view.getKernel1();
                check(mvd_get_kernel(s_, v, 1, k1.data.data()), "mvd_get_kernel");
                check(mvd_get_kernel(s_, v, 2, k2.data.data()), "mvd_get_kernel");
                view.kernel1_ = std::move(k1);
                view.kernel2_ = std::move(k2);
            }
            mvd_info info;
            check(mvd_get_info(s_, &info), "mvd_get_info");
            avg_ = info.avg;
            while (i_ < numIterations) runIteration();
            check(mvd_finish(s_), "mvd_finish");           // gen-2: "Masking never updated pixels."
        } catch (...) {
            mvd_session_destroy(s_);
            s_ = nullptr;
            throw;
        }
    }
    ~Deconvolution() override { if (s_) mvd_session_destroy(s_); }
    Deconvolution(const Deconvolution&) = delete;
    Deconvolution& operator=(const Deconvolution&) = delete;

    Input<V>& getData() { return views_; }
    std::string getName() const override { return name_; }
    double getAvg() const override { return avg_; }
    int getCurrentIteration() const { return i_; }
    const std::vector<ViewStat>& getStatistics() const { return stats_; }

    Image getPsi() override {
        const Image& first = views_.getViews()[0]->getImage();
        Image out(first.dims[0], first.dims[1], first.dims[2]);
        check(mvd_get_psi(s_, out.data.data()), "mvd_get_psi");
        return out;
    }
    void runIteration() override {
        const int V_ = views_.getNumViews();
        std::vector<double> s(V_), m(V_);
        check(mvd_run(s_, 1, s.data(), m.data()), "mvd_run");
        for (int v = 0; v < V_; ++v) stats_.push_back(ViewStat{i_, v, s[v], m[v]});
        ++i_;
    }

private:
    Input<V>& views_;
    std::string name_;
    mvd_session* s_ = nullptr;
    double avg_ = 0.0;
    int i_ = 0;
    std::vector<ViewStat> stats_;
};

using BayesMVDeconvolution = Deconvolution<LRFFT, 1>;
using MVDeconvolution = Deconvolution<MVDeconFFT, 2>;

}  // namespace spim_b200
