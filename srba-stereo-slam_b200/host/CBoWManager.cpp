#include "CBoWManager.h"
#include <zlib.h>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>

static bool field(const std::string& s, const char* key, std::string& out)
{
    const size_t p = s.find(key);
    if (p == std::string::npos) return false;
    size_t a = p + strlen(key), b = a;
    while (b < s.size() && s[b] != ',' && s[b] != '}' && s[b] != ' ' && s[b] != '\n' && s[b] != '\r') ++b;
    out = s.substr(a, b - a);
    return true;
}

// TemplatedVocabulary::load as restated in SURVEY.md A.3: node 0 = root, a node joins its parent's child list in
// file order, descriptor strings are MSB-first (char c <-> bit 255-c), DBoW2 bit i <-> byte i/8 bit i%8.
void loadFlatVocabulary(const std::string& filename, TFlatVocabulary& v)
{
    gzFile f = gzopen(filename.c_str(), "rb");          // zlib reads plain files transparently
    if (!f) throw std::runtime_error("cannot open vocabulary " + filename);
    std::vector<int> ids, parents; std::vector<double> weights; std::vector<std::string> descs; std::vector<std::pair<int, int>> words;
    std::vector<char> buf(1 << 16);
    bool pending = false; int pid = 0, nid = 0; double w = 0;
    std::string tmp;
    while (gzgets(f, buf.data(), (int)buf.size())) {
        std::string s(buf.data());
        if (s.find("- { nodeId:") != std::string::npos && s.find("wordId:") == std::string::npos) {
            field(s, "nodeId:", tmp); nid = atoi(tmp.c_str());
            field(s, "parentId:", tmp); pid = atoi(tmp.c_str());
            field(s, "weight:", tmp); w = atof(tmp.c_str());
            pending = true;
        }
        const size_t d = s.find("descriptor:");
        if (d != std::string::npos && pending) {
            const size_t q0 = s.find('"', d) + 1, q1 = s.find('"', q0);
            ids.push_back(nid); parents.push_back(pid); weights.push_back(w); descs.push_back(s.substr(q0, q1 - q0));
            pending = false;
        } else if (s.find("- { wordId:") != std::string::npos) {
            std::string a, b; field(s, "wordId:", a); field(s, "nodeId:", b);
            words.push_back(std::make_pair(atoi(a.c_str()), atoi(b.c_str())));
        } else {
            const char* keys[4] = { "k:", "L:", "scoringType:", "weightingType:" };
            int* dst[4] = { &v.k, &v.L, &v.scoring, &v.weighting };
            for (int i = 0; i < 4; ++i) {
                const size_t p = s.find_first_not_of(" ");
                if (p != std::string::npos && s.compare(p, strlen(keys[i]), keys[i]) == 0) *dst[i] = atoi(s.c_str() + p + strlen(keys[i]));
            }
        }
    }
    gzclose(f);
    if (ids.empty() || words.empty()) throw std::runtime_error("vocabulary " + filename + " has no nodes/words");
    int nn = 0; for (int id : ids) nn = std::max(nn, id + 1);
    v.nnodes = nn; v.nwords = (int)words.size();
    v.node_desc.assign((size_t)nn * 32, 0); v.node_weight.assign(nn, 0.0); v.node_word.assign(nn, -1);
    std::vector<int> cnt(nn, 0);
    for (int p : parents) cnt[p]++;
    v.child_ofs.assign(nn + 1, 0);
    for (int i = 0; i < nn; ++i) v.child_ofs[i + 1] = v.child_ofs[i] + cnt[i];
    std::vector<int> fill(v.child_ofs.begin(), v.child_ofs.end() - 1);
    v.child_idx.assign(ids.size(), 0);
    for (size_t i = 0; i < ids.size(); ++i) {
        v.child_idx[fill[parents[i]]++] = ids[i];
        v.node_weight[ids[i]] = weights[i];
        const std::string& bits = descs[i];
        if (bits.size() != 256) throw std::runtime_error("only 256-bit BRIEF vocabularies are supported");
        for (int c = 0; c < 256; ++c) if (bits[c] == '1') { const int bit = 255 - c; v.node_desc[(size_t)ids[i] * 32 + bit / 8] |= (uint8_t)(1u << (bit % 8)); }
    }
    for (const auto& wn : words) v.node_word[wn.second] = wn.first;
    if (v.scoring != 0 || v.weighting != 0) throw std::runtime_error("only L1 scoring + TF-IDF weighting (the shipped vocabulary) are supported");
}

CBoWManager::~CBoWManager() { if (m_bow) sfe_bow_destroy(m_bow); }

void CBoWManager::loadVocabulary(const std::string& voc_filename)
{
    m_host_voc.reset(new TFlatVocabulary()); loadFlatVocabulary(voc_filename, *m_host_voc);
    if (m_bow) { sfe_bow_destroy(m_bow); m_bow = nullptr; }
}

void CBoWManager::ensureUploaded()
{
    if (m_bow) return;
    if (!m_host_voc) throw std::runtime_error("CBoWManager: no vocabulary loaded");
    if (!m_ctx) m_ctx = rso::CStereoOdometryEstimator::currentContext();
    if (!m_ctx) throw std::runtime_error("CBoWManager: no GPU context yet (process a stereo pair first or call setContext)");
    const TFlatVocabulary& v = *m_host_voc;
    if (sfe_bow_create(m_ctx, v.nnodes, v.nwords, v.child_ofs.data(), v.child_idx.data(), v.node_desc.data(), v.node_weight.data(),
                       v.node_word.data(), &m_bow) != SFE_OK)
        throw std::runtime_error(std::string("sfe_bow_create: ") + sfe_last_error(m_ctx));
    m_host_voc.reset();
}

void CBoWManager::loadVocabulary(sfe_context* ctx, const std::string& voc_filename)
{
    m_ctx = ctx; loadVocabulary(voc_filename); ensureUploaded();
}

void CBoWManager::loadVocabularyFromConfigFile(const mrpt::utils::CConfigFile& config, const std::string& section, const std::string& param)
{
    std::string voc = config.read_string(section, param, "", true);
    if (voc.empty()) throw std::runtime_error("vocabulary file name is empty in [" + section + "] " + param);
    auto exists = [](const std::string& f) { std::ifstream t(f.c_str()); return t.good(); };
    if (!exists(voc)) {                                                // relative to the ini file's directory
        const std::string& ini = config.getAssociatedFile(); const size_t sl = ini.find_last_of('/');
        if (sl != std::string::npos && exists(ini.substr(0, sl + 1) + voc)) voc = ini.substr(0, sl + 1) + voc;
        else throw std::runtime_error("Assert file existence failed: " + voc);                   // ASSERT_FILE_EXISTS_, CBoWManager.h:63
    }
    loadVocabulary(voc);
}

void CBoWManager::loadVocabularyFromConfigFile(sfe_context* ctx, const std::string& ini_file, const std::string& section, const std::string& param)
{
    m_ctx = ctx; loadVocabularyFromConfigFile(mrpt::utils::CConfigFile(ini_file), section, param); ensureUploaded();
}

void CBoWManager::insertIntoDB(const CStereoSLAMKF& kf)
{
    ensureUploaded();
    int32_t id = 0;
    if (sfe_bow_db_add(m_bow, kf.m_descriptors_left.ptr(0), kf.m_descriptors_left.rows, &id) != SFE_OK)
        throw std::runtime_error(std::string("sfe_bow_db_add: ") + sfe_last_error(m_ctx));
}

void CBoWManager::queryDB(const CStereoSLAMKF& kf, QueryResults& ret, unsigned int num_results)
{
    ret.clear();
    ensureUploaded();
    std::vector<int32_t> ids(num_results); std::vector<double> sc(num_results); int32_t n = 0;
    if (sfe_bow_db_query(m_bow, kf.m_descriptors_left.ptr(0), kf.m_descriptors_left.rows, (int)num_results, ids.data(), sc.data(), &n) != SFE_OK)
        throw std::runtime_error(std::string("sfe_bow_db_query: ") + sfe_last_error(m_ctx));
    for (int i = 0; i < n; ++i) ret.push_back(DBoW2::Result((DBoW2::EntryId)ids[i], sc[i]));
}

int CBoWManager::size() const { return m_bow ? sfe_bow_db_size(m_bow) : 0; }

void matchKeyframes(sfe_context* ctx, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf, std::vector<cv::DMatch>& out)
{
    const size_t nq = this_kf.m_matches.size(), nt = other_kf.m_matches.size();
    std::vector<uint8_t> q(nq * 32), t(nt * 32);
    for (size_t k = 0; k < nq; ++k) std::memcpy(&q[k * 32], this_kf.m_descriptors_left.ptr(this_kf.m_matches[k].queryIdx), 32);     // EST.cpp:1354-1358
    for (size_t k = 0; k < nt; ++k) std::memcpy(&t[k * 32], other_kf.m_descriptors_left.ptr(other_kf.m_matches[k].queryIdx), 32);   // EST.cpp:1435-1439
    std::vector<sfe_dmatch> m(std::max<size_t>(nq, 1)); int32_t n = 0;
    if (sfe_match_bf(ctx, q.data(), (int)nq, t.data(), (int)nt, m.data(), &n) != SFE_OK)
        throw std::runtime_error(std::string("sfe_match_bf: ") + sfe_last_error(ctx));
    out.resize(n);
    if (n) std::memcpy((void*)out.data(), m.data(), sizeof(sfe_dmatch) * n);
}
