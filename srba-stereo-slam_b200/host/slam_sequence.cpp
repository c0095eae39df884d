// slam_sequence.cpp -- runs CSRBAStereoSLAMEstimator (stereo_slam_frontend.h) over a raw stereo sequence the way
// srba-stereo-slam_main.cpp:31-76 runs the reference over an image directory, and dumps what the SRBA engine would be
// fed: one observation list per keyframe (the argument of rba.define_new_keyframe, EST.cpp:166, 782), the keyframe-check
// decisions and the per-frame poses.  Used by tests/test_gpu_sequence.py (diffed against the oracle chain) and as the
// INTEGRATION.md example of driving the adapters without MRPT/SRBA.
//   slam_sequence <config.ini> <frames.bin> <W> <H> <nframes> <out.bin>
// frames.bin = nframes x {left, right} x H x W bytes.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include "stereo_slam_frontend.h"

struct RawSequence : public IStereoSource {
    std::vector<uint8_t> data; int W, H, n, next = 0;
    mrpt::obs::CObservationStereoImagesPtr getNextFrame() override
    {
        if (next >= n) return mrpt::obs::CObservationStereoImagesPtr();
        mrpt::obs::CObservationStereoImagesPtr o(new mrpt::obs::CObservationStereoImages());
        o->imageLeft.loadFromMemoryBuffer(W, H, false, &data[(size_t)(2 * next) * W * H]);
        o->imageRight.loadFromMemoryBuffer(W, H, false, &data[(size_t)(2 * next + 1) * W * H]);
        o->timestamp = (mrpt::system::TTimeStamp)next;
        ++next;
        return o;
    }
};
template <typename T> static void put(std::ofstream& o, const T& v) { o.write(reinterpret_cast<const char*>(&v), sizeof(T)); }

int main(int argc, char** argv)
{
    if (argc < 7) { fprintf(stderr, "usage: slam_sequence <config.ini> <frames.bin> <W> <H> <nframes> <out.bin>\n"); return 2; }
    try {
        mrpt::utils::CConfigFile config(argv[1]);
        general_options.loadFromConfigFile(config);                     // MAIN.cpp:46
        RawSequence src; src.W = atoi(argv[3]); src.H = atoi(argv[4]); src.n = atoi(argv[5]);
        src.data.resize((size_t)src.n * 2 * src.W * src.H);
        { std::ifstream f(argv[2], std::ios::binary); f.read(reinterpret_cast<char*>(src.data.data()), src.data.size()); if (!f) { fprintf(stderr, "short read\n"); return 2; } }
        CSRBAStereoSLAMEstimator est;
        est.initialize(config);                                          // MAIN.cpp:52
        est.setImageSource(&src);
        est.performStereoSLAM();                                         // MAIN.cpp:55
        std::ofstream out(argv[6], std::ios::binary);
        const auto& lists = est.observationLists(); const auto& kff = est.keyframeFrames();
        put(out, (int32_t)lists.size());
        for (size_t k = 0; k < lists.size(); ++k) {
            put(out, (int32_t)kff[k]); put(out, (int32_t)lists[k].size());
            for (const TObservation& o : lists[k]) { put(out, (int64_t)o.feat_id); put(out, o.lx); put(out, o.ly); put(out, o.rx); put(out, o.ry); put(out, o.X); put(out, o.Y); put(out, o.Z); }
        }
        const auto& checks = est.checks();
        put(out, (int32_t)checks.size());
        for (const TKfCheckRecord& c : checks) {
            put(out, (int32_t)c.frame); put(out, (int32_t)c.candidate_id);
            put(out, (int32_t)(c.vo_force_new_kf | (c.vo_force_check_tracking << 1) | (c.vo_force_check_distance << 2) | (c.potential_lc << 3) | (c.confirmed_lc << 4) | (c.inserted << 5)));
            put(out, (int32_t)c.lc_id); put(out, (int32_t)c.n_tracked_ids); put(out, (int32_t)c.n_new_ids);
            put(out, (int32_t)c.query.size()); for (const auto& q : c.query) { put(out, (int32_t)q.first); put(out, q.second); }
            put(out, (int32_t)c.similar_kfs.size()); for (TKeyFrameID s : c.similar_kfs) put(out, (int32_t)s);
            put(out, (int32_t)c.tracked.size()); for (const auto& t : c.tracked) { put(out, (int32_t)t.first); put(out, (int32_t)t.second); }
        }
        const auto& poses = est.framePoses();
        put(out, (int32_t)poses.size());
        for (const auto& p : poses) for (int i = 0; i < 6; ++i) put(out, p[i]);
        std::cout << "slam_sequence ok: " << est.framesProcessed() << " frames, " << lists.size() << " keyframes, " << checks.size() << " checks" << std::endl;
    } catch (std::exception& e) { std::cerr << "EXCEPTION: " << e.what() << std::endl; return 1; }
    return 0;
}
