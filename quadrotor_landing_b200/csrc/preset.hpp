// preset.hpp -- host-side loader for the node's parameter files.
//
// The reference loads its estimator parameters from a flat YAML file through the ROS parameter server
// (quad_state_estimation/config/relative_pose_EKF_{rotors,hardware}.yaml, read in
// src/relative_pose_EKF_node.cpp:20-136).  This is the ROS-free equivalent: the same keys, the same
// defaults for missing scalars (the `node.param<>` defaults), the same conventions (q_vc in x,y,z,w array
// order, node.cpp:108-110; camera_K row-major, :115-117; three numbers per tag position, :128-136).
// Only the subset of YAML those files use is understood: `key: scalar`, `key: [a, b, ...]` (the list may run
// over several lines), `#` comments, quoted strings, True/False.
#pragma once

#include <cctype>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "../../include/qekf.h"

namespace qekf {
namespace preset {

struct Doc {
    std::map<std::string, std::string> scalars;
    std::map<std::string, std::vector<double>> lists;
};

inline std::string trim(const std::string &s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

// strip a trailing comment (a '#' outside quotes)
inline std::string strip_comment(const std::string &line)
{
    bool in_s = false, in_d = false;
    for (size_t i = 0; i < line.size(); ++i) {
        const char ch = line[i];
        if (ch == '\'' && !in_d) in_s = !in_s;
        else if (ch == '"' && !in_s) in_d = !in_d;
        else if (ch == '#' && !in_s && !in_d) return line.substr(0, i);
    }
    return line;
}

inline bool parse_number(const std::string &tok, double *out)
{
    const std::string t = trim(tok);
    if (t.empty()) return false;
    char *end = nullptr;
    const double v = std::strtod(t.c_str(), &end);
    if (end == t.c_str() || *end != '\0') return false;
    *out = v;
    return true;
}

// returns an empty string on success, otherwise a description of the first problem
inline std::string parse(const std::string &text, Doc *doc)
{
    size_t pos = 0;
    int line_no = 0;
    std::string pending_key, pending_list;   // a list whose closing bracket has not been seen yet
    int pending_line = 0;
    auto finish_list = [&](const std::string &key, const std::string &body, int at) -> std::string {
        std::vector<double> vals;
        size_t p = 0;
        while (p <= body.size()) {
            size_t q = body.find(',', p);
            if (q == std::string::npos) q = body.size();
            const std::string tok = trim(body.substr(p, q - p));
            if (!tok.empty()) {
                double v;
                if (!parse_number(tok, &v)) return "line " + std::to_string(at) + ": '" + tok + "' in list '" + key + "' is not a number";
                vals.push_back(v);
            }
            p = q + 1;
        }
        doc->lists[key] = vals;
        return "";
    };
    while (pos <= text.size()) {
        size_t eol = text.find('\n', pos);
        if (eol == std::string::npos) eol = text.size();
        ++line_no;
        const std::string line = trim(strip_comment(text.substr(pos, eol - pos)));
        pos = eol + 1;
        if (line.empty()) continue;
        if (!pending_key.empty()) {
            const size_t close = line.find(']');
            pending_list += " " + (close == std::string::npos ? line : line.substr(0, close));
            if (close != std::string::npos) {
                const std::string err = finish_list(pending_key, pending_list, pending_line);
                if (!err.empty()) return err;
                pending_key.clear();
                pending_list.clear();
            }
            continue;
        }
        const size_t colon = line.find(':');
        if (colon == std::string::npos) return "line " + std::to_string(line_no) + ": expected 'key: value'";
        const std::string key = trim(line.substr(0, colon));
        std::string val = trim(line.substr(colon + 1));
        if (key.empty()) return "line " + std::to_string(line_no) + ": empty key";
        if (!val.empty() && val[0] == '[') {
            const size_t close = val.find(']');
            if (close == std::string::npos) {
                pending_key = key;
                pending_list = val.substr(1);
                pending_line = line_no;
            } else {
                const std::string err = finish_list(key, val.substr(1, close - 1), line_no);
                if (!err.empty()) return err;
            }
            continue;
        }
        if (val.size() >= 2 && (val.front() == '"' || val.front() == '\'') && val.back() == val.front())
            val = val.substr(1, val.size() - 2);
        doc->scalars[key] = val;
    }
    if (!pending_key.empty()) return "list '" + pending_key + "' (line " + std::to_string(pending_line) + ") is never closed";
    return "";
}

inline bool as_bool(const std::string &v, bool *out)
{
    std::string t;
    for (char ch : v) t.push_back((char)std::tolower((unsigned char)ch));
    if (t == "true" || t == "yes" || t == "on" || t == "1") { *out = true; return true; }
    if (t == "false" || t == "no" || t == "off" || t == "0") { *out = false; return true; }
    return false;
}

// Fill *p the way RelativePoseEKFNode's constructor does (node.cpp:20-136).  On entry *p must hold the class
// defaults (qekf_default_params); keys the file does not mention keep the node's `param<>` default.
inline std::string apply(const Doc &d, qekf_params *p)
{
    std::string err;
    auto num = [&](const char *key, double *dst, double node_default) {
        *dst = node_default;
        auto it = d.scalars.find(key);
        if (it == d.scalars.end()) return;
        double v;
        if (!parse_number(it->second, &v)) { if (err.empty()) err = std::string("'") + key + "' is not a number"; return; }
        *dst = v;
    };
    auto flag = [&](const char *key, int32_t *dst, bool node_default) {
        *dst = node_default;
        auto it = d.scalars.find(key);
        if (it == d.scalars.end()) return;
        bool b;
        if (!as_bool(it->second, &b)) { if (err.empty()) err = std::string("'") + key + "' is not a boolean"; return; }
        *dst = b;
    };
    // getParam semantics: the member keeps its previous value when the key is missing
    auto vec = [&](const char *key, double *dst, size_t n) {
        auto it = d.lists.find(key);
        if (it == d.lists.end()) return;
        if (it->second.size() != n) { if (err.empty()) err = std::string("'") + key + "' needs " + std::to_string(n) + " numbers"; return; }
        for (size_t i = 0; i < n; ++i) dst[i] = it->second[i];
    };
    num("update_freq", &p->update_freq, 100.0);                                   // node.cpp:35-40
    num("measurement_freq", &p->measurement_freq, 10.0);
    num("measurement_delay", &p->measurement_delay, 0.010);
    num("measurement_delay_max", &p->measurement_delay_max, 0.200);
    num("dyn_measurement_delay_offset", &p->dyn_measurement_delay_offset, 0.0);
    flag("limit_measurement_freq", &p->limit_measurement_freq, false);
    flag("est_bias", &p->est_bias, true);                                         // node.cpp:60-64
    flag("corner_margin_enbl", &p->corner_margin_enbl, true);
    flag("direct_orien_method", &p->direct_orien_method, false);
    flag("multirate_ekf", &p->multirate_ekf, false);
    flag("dynamic_meas_delay", &p->dynamic_meas_delay, false);
    vec("Q_a_diag", p->Q_a, 3); vec("Q_w_diag", p->Q_w, 3);                       // node.cpp:66-79
    vec("Q_ab_diag", p->Q_ab, 3); vec("Q_wb_diag", p->Q_wb, 3);
    vec("R_r_diag", p->R_r, 3); vec("R_ang_diag", p->R_ang, 3);                   // node.cpp:81-87
    num("r_cov_init", &p->r_cov_init, 0.1);                                       // node.cpp:89-93
    num("v_cov_init", &p->v_cov_init, 0.1);
    num("ang_cov_init", &p->ang_cov_init, 0.15);
    num("ab_cov_init", &p->ab_cov_init, 0.5);
    num("wb_cov_init", &p->wb_cov_init, 0.1);
    vec("accel_bias_static", p->ab_static, 3);                                    // node.cpp:95-101
    vec("gyro_bias_static", p->wb_static, 3);
    vec("r_v_cv", p->r_v_cv, 3);                                                  // node.cpp:103-110
    vec("q_vc", p->q_vc, 4);                                                      // x,y,z,w array order
    {
        double w = p->camera_width, h = p->camera_height;                         // node.cpp:112-113
        num("camera_width", &w, w);
        num("camera_height", &h, h);
        p->camera_width = (int32_t)w;
        p->camera_height = (int32_t)h;
    }
    vec("camera_K", p->camera_K, 9);                                              // row-major, node.cpp:115-117
    {
        double n = p->n_tags;                                                     // node.cpp:119-120
        num("n_tags", &n, n);
        if (n < 0 || n > QEKF_MAX_TAGS) { if (err.empty()) err = "n_tags out of range (0.." + std::to_string(QEKF_MAX_TAGS) + ")"; }
        else p->n_tags = (int32_t)n;
    }
    num("tag_in_view_margin", &p->tag_in_view_margin, p->tag_in_view_margin);
    if (err.empty()) {
        auto w = d.lists.find("tag_widths");                                      // node.cpp:122-136
        auto q = d.lists.find("tag_positions");
        if (w != d.lists.end()) {
            if ((int)w->second.size() < p->n_tags) err = "'tag_widths' has fewer entries than n_tags";
            else for (int i = 0; i < p->n_tags; ++i) p->tag_widths[i] = w->second[(size_t)i];
        }
        if (err.empty() && q != d.lists.end()) {
            if ((int)q->second.size() < 3 * p->n_tags) err = "'tag_positions' needs 3 numbers per tag";
            else for (int i = 0; i < 3 * p->n_tags; ++i) p->tag_positions[i] = q->second[(size_t)i];
        }
    }
    return err;
}

}  // namespace preset
}  // namespace qekf
