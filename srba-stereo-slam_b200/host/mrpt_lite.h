// mrpt_lite.h -- the handful of MRPT (>= 1.3) value types the reference's hot-path callers hand to the stereo front-end,
// for builds where MRPT itself is not installed (this container, the GPU box).  Define SRBA_B200_WITH_MRPT to use the
// real headers instead; the names, namespaces and member signatures below are MRPT's, so the reference's call lines
// (CSRBAStereoSLAMEstimator.cpp:111-114, 263-330, 1102-1144; tests/conformance/ref_calls.cpp) compile against either.
//
//   mrpt::utils::CConfigFile            .ini reader (read_string/int/double/float/bool/uint64_t/read_vector)
//   mrpt::poses::CPose3DRotVec          SE(3) pose as translation + rotation vector (m_coords, m_rotvec), composition
//   mrpt::poses::CPose3D                SE(3) pose as translation + yaw/pitch/roll
//   mrpt::utils::TCamera/TStereoCamera  pinhole intrinsics + left-to-right pose, loadFromConfigFile("CAMERA", cfg)
//   mrpt::utils::CImage                 8-bit grey image view (getWidth/getHeight + raw rows)
//   mrpt::obs::CObservationStereoImages(+Ptr)   what CCameraSensor::getNextFrame() yields (EST.cpp:44,79)
#pragma once
#ifdef SRBA_B200_WITH_MRPT
#include <mrpt/obs/CObservationStereoImages.h>
#include <mrpt/poses/CPose3D.h>
#include <mrpt/poses/CPose3DRotVec.h>
#include <mrpt/utils/CConfigFile.h>
#include <mrpt/utils/TStereoCamera.h>
#else
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef DEG2RAD
#define DEG2RAD(x) ((x) * M_PI / 180.0)
#define RAD2DEG(x) ((x) * 180.0 / M_PI)
#endif
#ifndef THROW_EXCEPTION
#define THROW_EXCEPTION(msg) throw std::logic_error(std::string(msg));
#define ASSERT_(cond) if (!(cond)) throw std::logic_error(std::string("Assert condition failed: ") + #cond);
#define ASSERTDEB_(cond) ASSERT_(cond)
#define MRPT_TRY_START try {
#define MRPT_TRY_END } catch (std::exception& e__) { throw std::logic_error(std::string(e__.what()) + "\n" + __FILE__ ":" + std::to_string(__LINE__)); }
#define MRPT_LOAD_CONFIG_VAR(var, type, cfg, sec) var = cfg.read_##type(sec, #var, var);
#define SHOW_WARNING(msg) std::cerr << "[WARNING] " << msg << std::endl;
#endif

namespace mrpt {
inline std::string format(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
}
#include <cstdarg>
inline std::string mrpt::format(const char* fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    return std::string(buf);
}

namespace mrpt {
namespace system { typedef uint64_t TTimeStamp; }

namespace utils {
// ---- .ini file -----------------------------------------------------------------------------------------------------
// MRPT's CConfigFile on SimpleIni: [SECTION], `key = value`, `//` and `;` start comments, section and key names are
// case-insensitive.  read_*(section, name, default, failIfNotFound = false).
class CConfigFile {
public:
    CConfigFile() {}
    explicit CConfigFile(const std::string& fileName) { setFileName(fileName); }
    void setFileName(const std::string& fileName)
    {
        m_file = fileName; m_values.clear();
        std::ifstream f(fileName.c_str());
        if (!f) throw std::runtime_error("CConfigFile: cannot open " + fileName);
        std::string line, sec;
        while (std::getline(f, line)) {
            const size_t c = line.find("//"); if (c != std::string::npos) line.erase(c);
            line = trim(line);
            if (line.empty() || line[0] == ';' || line[0] == '#') continue;
            if (line[0] == '[') { sec = upper(trim(line.substr(1, line.find(']') - 1))); continue; }
            const size_t eq = line.find('=');
            if (eq == std::string::npos) continue;
            m_values[sec][upper(trim(line.substr(0, eq)))] = trim(line.substr(eq + 1));
        }
    }
    const std::string& getAssociatedFile() const { return m_file; }
    bool sectionExists(const std::string& section) const { return m_values.count(upper(section)) != 0; }

    std::string read_string(const std::string& section, const std::string& name, const std::string& def, bool failIfNotFound = false) const
    {
        std::string v; return lookup(section, name, v, failIfNotFound) ? v : def;
    }
    int read_int(const std::string& section, const std::string& name, int def, bool failIfNotFound = false) const
    {
        std::string v; return lookup(section, name, v, failIfNotFound) ? (int)strtol(v.c_str(), nullptr, 0) : def;
    }
    uint64_t read_uint64_t(const std::string& section, const std::string& name, uint64_t def, bool failIfNotFound = false) const
    {
        std::string v; return lookup(section, name, v, failIfNotFound) ? (uint64_t)strtoull(v.c_str(), nullptr, 0) : def;
    }
    double read_double(const std::string& section, const std::string& name, double def, bool failIfNotFound = false) const
    {
        std::string v; return lookup(section, name, v, failIfNotFound) ? atof(v.c_str()) : def;
    }
    float read_float(const std::string& section, const std::string& name, float def, bool failIfNotFound = false) const
    {
        return (float)read_double(section, name, def, failIfNotFound);
    }
    bool read_bool(const std::string& section, const std::string& name, bool def, bool failIfNotFound = false) const
    {
        std::string v; if (!lookup(section, name, v, failIfNotFound)) return def;
        v = upper(v);
        return v == "TRUE" || v == "YES" || v == "T" || v == "Y" || atoi(v.c_str()) != 0;
    }
    // "[a b c]" / "[a, b, c]"
    template <class VEC> void read_vector(const std::string& section, const std::string& name, const VEC& def, VEC& out, bool failIfNotFound = false) const
    {
        std::string v; if (!lookup(section, name, v, failIfNotFound)) { out = def; return; }
        for (char& ch : v) if (ch == '[' || ch == ']' || ch == ',' || ch == ';') ch = ' ';
        std::istringstream is(v); out.clear();
        double x; while (is >> x) out.push_back((typename VEC::value_type)x);
    }

private:
    static std::string trim(const std::string& s)
    {
        const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }
    static std::string upper(std::string s) { for (char& c : s) c = (char)toupper((unsigned char)c); return s; }
    bool lookup(const std::string& section, const std::string& name, std::string& out, bool fail) const
    {
        auto s = m_values.find(upper(section));
        if (s != m_values.end()) { auto k = s->second.find(upper(name)); if (k != s->second.end()) { out = k->second; return true; } }
        if (fail) throw std::runtime_error("Value '" + name + "' not found in section '" + section + "' of file '" + m_file + "'");
        return false;
    }
    std::string m_file;
    std::map<std::string, std::map<std::string, std::string>> m_values;
};
}  // namespace utils

namespace math {
struct TPoint3D { double x, y, z; TPoint3D(double x_ = 0, double y_ = 0, double z_ = 0) : x(x_), y(y_), z(z_) {} };
// Fixed-size column vector with the two Eigen members the reference uses (operator[], norm()).
template <int N> struct CArrayDouble {
    double v[N];
    CArrayDouble() { for (int i = 0; i < N; ++i) v[i] = 0; }
    double& operator[](int i) { return v[i]; }
    const double& operator[](int i) const { return v[i]; }
    double norm() const { double s = 0; for (int i = 0; i < N; ++i) s += v[i] * v[i]; return std::sqrt(s); }
    static int size() { return N; }
};
}  // namespace math

namespace poses {
struct TRot3 { double m[3][3]; };
inline TRot3 rot_mul(const TRot3& a, const TRot3& b)
{
    TRot3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += a.m[i][k] * b.m[k][j]; r.m[i][j] = s; }
    return r;
}
// exp / log on SO(3) (Rodrigues)
inline TRot3 rot_exp(const math::CArrayDouble<3>& w)
{
    const double th = w.norm(); TRot3 R;
    double a, b;                                      // R = I + a [w]x + b [w]x^2
    if (th < 1e-8) { a = 1.0 - th * th / 6.0; b = 0.5 - th * th / 24.0; } else { a = std::sin(th) / th; b = (1.0 - std::cos(th)) / (th * th); }
    const double x = w[0], y = w[1], z = w[2];
    R.m[0][0] = 1 - b * (y * y + z * z); R.m[0][1] = -a * z + b * x * y;      R.m[0][2] = a * y + b * x * z;
    R.m[1][0] = a * z + b * x * y;       R.m[1][1] = 1 - b * (x * x + z * z); R.m[1][2] = -a * x + b * y * z;
    R.m[2][0] = -a * y + b * x * z;      R.m[2][1] = a * x + b * y * z;       R.m[2][2] = 1 - b * (x * x + y * y);
    return R;
}
inline math::CArrayDouble<3> rot_log(const TRot3& R)
{
    math::CArrayDouble<3> w;
    const double tr = R.m[0][0] + R.m[1][1] + R.m[2][2];
    double c = 0.5 * (tr - 1.0); c = std::max(-1.0, std::min(1.0, c));
    const double th = std::acos(c);
    const double ax = R.m[2][1] - R.m[1][2], ay = R.m[0][2] - R.m[2][0], az = R.m[1][0] - R.m[0][1];
    if (th < 1e-8) { w[0] = 0.5 * ax; w[1] = 0.5 * ay; w[2] = 0.5 * az; return w; }
    if (M_PI - th < 1e-6) {                           // near pi: axis from the symmetric part
        double d[3] = { R.m[0][0], R.m[1][1], R.m[2][2] }; int i = d[0] > d[1] ? (d[0] > d[2] ? 0 : 2) : (d[1] > d[2] ? 1 : 2);
        double u[3]; u[i] = std::sqrt(std::max(0.0, 0.5 * (d[i] + 1.0)));
        for (int j = 0; j < 3; ++j) if (j != i) u[j] = (R.m[i][j] + R.m[j][i]) / (4.0 * u[i]);
        const double sgn = (ax * u[0] + ay * u[1] + az * u[2]) < 0 ? -1.0 : 1.0;
        for (int j = 0; j < 3; ++j) w[j] = sgn * th * u[j];
        return w;
    }
    const double k = th / (2.0 * std::sin(th));
    w[0] = k * ax; w[1] = k * ay; w[2] = k * az;
    return w;
}

class CPose3DRotVec;
// Translation + yaw (Z) / pitch (Y) / roll (X), R = Rz(yaw) Ry(pitch) Rx(roll) -- MRPT's CPose3D convention.
class CPose3D {
public:
    CPose3D() { setFromValues(0, 0, 0, 0, 0, 0); }
    CPose3D(double x, double y, double z, double yaw = 0, double pitch = 0, double roll = 0) { setFromValues(x, y, z, yaw, pitch, roll); }
    CPose3D(const CPose3DRotVec& p);            // implicit, as the reference relies on (EST.cpp:1765, 1850)
    void setFromValues(double x, double y, double z, double yaw, double pitch, double roll)
    {
        m_coords[0] = x; m_coords[1] = y; m_coords[2] = z;
        const double cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch), cr = std::cos(roll), sr = std::sin(roll);
        m_R.m[0][0] = cy * cp; m_R.m[0][1] = cy * sp * sr - sy * cr; m_R.m[0][2] = cy * sp * cr + sy * sr;
        m_R.m[1][0] = sy * cp; m_R.m[1][1] = sy * sp * sr + cy * cr; m_R.m[1][2] = sy * sp * cr - cy * sr;
        m_R.m[2][0] = -sp;     m_R.m[2][1] = cp * sr;                m_R.m[2][2] = cp * cr;
    }
    void composeFrom(const CPose3D& A, const CPose3D& B)          // this = A (+) B
    {
        const TRot3 R = rot_mul(A.m_R, B.m_R); math::CArrayDouble<3> t;
        for (int i = 0; i < 3; ++i) t[i] = A.m_coords[i] + A.m_R.m[i][0] * B.m_coords[0] + A.m_R.m[i][1] * B.m_coords[1] + A.m_R.m[i][2] * B.m_coords[2];
        m_R = R; m_coords = t;
    }
    CPose3D operator+(const CPose3D& b) const { CPose3D r; r.composeFrom(*this, b); return r; }
    double x() const { return m_coords[0]; } double y() const { return m_coords[1]; } double z() const { return m_coords[2]; }
    void getYawPitchRoll(double& yaw, double& pitch, double& roll) const
    {
        pitch = std::atan2(-m_R.m[2][0], std::hypot(m_R.m[0][0], m_R.m[1][0]));
        if (std::fabs(std::fabs(pitch) - M_PI / 2) < 1e-9) { yaw = std::atan2(-m_R.m[0][1], m_R.m[1][1]); roll = 0; }
        else { yaw = std::atan2(m_R.m[1][0], m_R.m[0][0]); roll = std::atan2(m_R.m[2][1], m_R.m[2][2]); }
    }
    double yaw() const { double y, p, r; getYawPitchRoll(y, p, r); return y; }
    double pitch() const { double y, p, r; getYawPitchRoll(y, p, r); return p; }
    double roll() const { double y, p, r; getYawPitchRoll(y, p, r); return r; }
    const TRot3& getRotationMatrix() const { return m_R; }
    math::CArrayDouble<3> m_coords; TRot3 m_R;
};

// Translation m_coords + rotation vector m_rotvec (axis * angle).  Composition follows SE(3): (A + B) maps B's frame
// through A.  Members and operators are MRPT's (EST.cpp:327-330, 367-368, 1112-1114, 1765, 1850-1866, UTL.h:558-574).
class CPose3DRotVec {
public:
    math::CArrayDouble<3> m_coords, m_rotvec;
    CPose3DRotVec() {}
    CPose3DRotVec(double w1, double w2, double w3, double x, double y, double z) { m_rotvec[0] = w1; m_rotvec[1] = w2; m_rotvec[2] = w3; m_coords[0] = x; m_coords[1] = y; m_coords[2] = z; }
    explicit CPose3DRotVec(const math::CArrayDouble<6>& v) { for (int i = 0; i < 3; ++i) { m_rotvec[i] = v[i]; m_coords[i] = v[3 + i]; } }   // [w1 w2 w3 x y z]
    explicit CPose3DRotVec(const std::vector<double>& v) { for (int i = 0; i < 3 && v.size() >= 6; ++i) { m_rotvec[i] = v[i]; m_coords[i] = v[3 + i]; } }
    explicit CPose3DRotVec(const CPose3D& p) { m_coords = p.m_coords; m_rotvec = rot_log(p.getRotationMatrix()); }
    TRot3 getRotationMatrix() const { return rot_exp(m_rotvec); }
    void composeFrom(const CPose3DRotVec& A, const CPose3DRotVec& B)
    {
        const TRot3 RA = A.getRotationMatrix(), R = rot_mul(RA, B.getRotationMatrix()); math::CArrayDouble<3> t;
        for (int i = 0; i < 3; ++i) t[i] = A.m_coords[i] + RA.m[i][0] * B.m_coords[0] + RA.m[i][1] * B.m_coords[1] + RA.m[i][2] * B.m_coords[2];
        m_coords = t; m_rotvec = rot_log(R);
    }
    CPose3DRotVec operator+(const CPose3DRotVec& b) const { CPose3DRotVec r; r.composeFrom(*this, b); return r; }
    CPose3DRotVec& operator+=(const CPose3DRotVec& b) { composeFrom(*this, b); return *this; }
    math::TPoint3D operator+(const math::TPoint3D& p) const          // the point as seen from this pose's reference frame
    {
        const TRot3 R = getRotationMatrix();
        return math::TPoint3D(m_coords[0] + R.m[0][0] * p.x + R.m[0][1] * p.y + R.m[0][2] * p.z,
                              m_coords[1] + R.m[1][0] * p.x + R.m[1][1] * p.y + R.m[1][2] * p.z,
                              m_coords[2] + R.m[2][0] * p.x + R.m[2][1] * p.y + R.m[2][2] * p.z);
    }
    void inverse()
    {
        const TRot3 R = getRotationMatrix(); math::CArrayDouble<3> t;
        for (int i = 0; i < 3; ++i) t[i] = -(R.m[0][i] * m_coords[0] + R.m[1][i] * m_coords[1] + R.m[2][i] * m_coords[2]);
        m_coords = t; for (int i = 0; i < 3; ++i) m_rotvec[i] = -m_rotvec[i];
    }
    CPose3DRotVec getInverse() const { CPose3DRotVec r(*this); r.inverse(); return r; }
    double& operator[](int i) { return i < 3 ? m_coords[i] : m_rotvec[i - 3]; }       // MRPT order: x y z wx wy wz
    const double& operator[](int i) const { return i < 3 ? m_coords[i] : m_rotvec[i - 3]; }
};
inline CPose3D::CPose3D(const CPose3DRotVec& p) { m_coords = p.m_coords; m_R = p.getRotationMatrix(); }
inline std::ostream& operator<<(std::ostream& o, const CPose3DRotVec& p)
{
    return o << "(x,y,z,wx,wy,wz)=(" << p.m_coords[0] << "," << p.m_coords[1] << "," << p.m_coords[2] << "," << p.m_rotvec[0] << "," << p.m_rotvec[1] << "," << p.m_rotvec[2] << ")";
}
inline std::ostream& operator<<(std::ostream& o, const CPose3D& p)
{
    double y, pi, r; p.getYawPitchRoll(y, pi, r);
    return o << "(x,y,z,yaw,pitch,roll)=(" << p.x() << "," << p.y() << "," << p.z() << "," << RAD2DEG(y) << "deg," << RAD2DEG(pi) << "deg," << RAD2DEG(r) << "deg)";
}
// x y z qr qx qy qz, indexable like MRPT's CPose3DQuat (the reference reads rightCameraPose[0] = baseline, UTL.h:569)
struct CPose3DQuat { double v[7]; CPose3DQuat() : v{ 0, 0, 0, 1, 0, 0, 0 } {} double& operator[](int i) { return v[i]; } const double& operator[](int i) const { return v[i]; } };
}  // namespace poses

namespace utils {
struct TCamera {
    uint32_t ncols = 640, nrows = 480;
    double m_fx = 0, m_fy = 0, m_cx = 0, m_cy = 0; std::vector<double> dist = std::vector<double>(5, 0.0);
    double cx() const { return m_cx; } double cy() const { return m_cy; } double fx() const { return m_fx; } double fy() const { return m_fy; }
    void setIntrinsicParamsFromValues(double fx, double fy, double cx, double cy) { m_fx = fx; m_fy = fy; m_cx = cx; m_cy = cy; }
    // [<section>]: resolution = [ncols nrows], cx, cy, fx, fy, dist = [k1 k2 t1 t2 k3]   (kitti.ini:138-152)
    void loadFromConfigFile(const std::string& section, const CConfigFile& cfg)
    {
        std::vector<double> res; cfg.read_vector(section, "resolution", std::vector<double>(), res, true);
        if (res.size() != 2) throw std::runtime_error("TCamera: 'resolution' must have 2 elements in [" + section + "]");
        ncols = (uint32_t)res[0]; nrows = (uint32_t)res[1];
        m_fx = cfg.read_double(section, "fx", 0, true); m_fy = cfg.read_double(section, "fy", 0, true);
        m_cx = cfg.read_double(section, "cx", 0, true); m_cy = cfg.read_double(section, "cy", 0, true);
        cfg.read_vector(section, "dist", std::vector<double>(5, 0.0), dist, false);
    }
};
struct TStereoCamera {
    TCamera leftCamera, rightCamera; poses::CPose3DQuat rightCameraPose;
    // sections <section>_LEFT, <section>_RIGHT, <section>_LEFT2RIGHT_POSE (pose_quaternion = [x y z qr qx qy qz])
    void loadFromConfigFile(const std::string& section, const CConfigFile& cfg)
    {
        leftCamera.loadFromConfigFile(section + "_LEFT", cfg); rightCamera.loadFromConfigFile(section + "_RIGHT", cfg);
        std::vector<double> q; cfg.read_vector(section + "_LEFT2RIGHT_POSE", "pose_quaternion", std::vector<double>(), q, true);
        if (q.size() != 7) throw std::runtime_error("TStereoCamera: pose_quaternion must have 7 elements");
        for (int i = 0; i < 7; ++i) rightCameraPose[i] = q[i];
    }
};

// 8-bit grey image: a view over caller memory (loadFromMemoryBuffer, MRPT's raw-pixels loader) or an owned copy.
class CImage {
public:
    CImage() {}
    void loadFromMemoryBuffer(unsigned int width, unsigned int height, bool color, unsigned char* rawpixels, bool /*swapRedBlue*/ = false)
    {
        if (color) throw std::logic_error("CImage (lite): only 8-bit grey images");
        m_w = width; m_h = height; m_stride = width; m_data = rawpixels; m_own.reset();
    }
    void setFromGreyView(const unsigned char* p, unsigned int width, unsigned int height, size_t stride) { m_w = width; m_h = height; m_stride = stride; m_data = p; m_own.reset(); }
    void copyFrom(const unsigned char* p, unsigned int width, unsigned int height, size_t stride)
    {
        m_own.reset(new std::vector<unsigned char>((size_t)width * height));
        for (unsigned int r = 0; r < height; ++r) std::copy(p + r * stride, p + r * stride + width, m_own->begin() + (size_t)r * width);
        m_w = width; m_h = height; m_stride = width; m_data = m_own->data();
    }
    size_t getWidth() const { return m_w; } size_t getHeight() const { return m_h; } size_t getRowStride() const { return m_stride; }
    bool isColor() const { return false; }
    const unsigned char* get_unsafe(unsigned int col, unsigned int row, unsigned int /*channel*/ = 0) const { return m_data + (size_t)row * m_stride + col; }
private:
    unsigned int m_w = 0, m_h = 0; size_t m_stride = 0; const unsigned char* m_data = nullptr;
    std::shared_ptr<std::vector<unsigned char>> m_own;
};
}  // namespace utils

namespace obs {
struct CObservation { system::TTimeStamp timestamp = 0; virtual ~CObservation() {} };
struct CObservationStereoImages : public CObservation { utils::CImage imageLeft, imageRight; };
// MRPT 1.x smart pointers: shared ownership + present()
template <class T> struct TSmartPtr : public std::shared_ptr<T> {
    TSmartPtr() {}
    TSmartPtr(T* p) : std::shared_ptr<T>(p) {}
    template <class U> TSmartPtr(const std::shared_ptr<U>& o) : std::shared_ptr<T>(std::static_pointer_cast<T>(o)) {}
    bool present() const { return this->get() != nullptr; }
    void clear() { this->reset(); }
};
typedef TSmartPtr<CObservation> CObservationPtr;
typedef TSmartPtr<CObservationStereoImages> CObservationStereoImagesPtr;
}  // namespace obs
}  // namespace mrpt
#endif  // SRBA_B200_WITH_MRPT
