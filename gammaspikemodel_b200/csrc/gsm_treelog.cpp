// gsm_treelog.cpp — the step after the hot path: the tree-log line of one state.
//
// StumpedTreeLogger.toNewick (logger/StumpedTreeLogger.java:77-199; BEAST core TreeWithMetaDataLogger writes the same
// shape) walks the tree and prints, per node, `[&nstubs=..,rate=..,<id>=..,...]:<length>` with every number formatted by
// Java's StringBuffer.append(double) = Double.toString.  The per-node values are what the device already produced
// (gsm_eval: rates, weighted spikes; gsm_stub_choose: sampled stub counts); this file turns them into the bytes of the
// `.trees` file on the host.  Plain C++17, no CUDA.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gsm_b200.h"

namespace {

// Double.toString(double) as specified since JDK 19 (the shortest decimal that rounds to the double; earlier JDKs print a
// few values with one digit more): "NaN", "Infinity", "-Infinity", "0.0", "-0.0"; decimal notation with at least one
// fraction digit for 1e-3 <= |d| < 1e7; otherwise computerised scientific notation d.dddE[-]n.
void java_double(double d, std::string& out) {
    if (std::isnan(d)) { out += "NaN"; return; }
    if (std::isinf(d)) { out += d < 0 ? "-Infinity" : "Infinity"; return; }
    if (d == 0.0) { out += std::signbit(d) ? "-0.0" : "0.0"; return; }
    char buf[40];
    auto r = std::to_chars(buf, buf + sizeof buf, d, std::chars_format::scientific);   // shortest round-trip digits
    std::string s(buf, r.ptr);
    const bool neg = s[0] == '-';
    if (neg) s.erase(0, 1);
    const size_t e = s.find('e');
    std::string digits = s.substr(0, e);                 // "d" or "d.ddd"
    const int exp10 = std::stoi(s.substr(e + 1));
    std::string dg;
    for (char c : digits) if (c != '.') dg += c;         // significant digits, first one non-zero
    if (dg.size() == 1) {
        // Java renders at least two digits and takes the two-digit decimal closest to the value (Double.MIN_VALUE is
        // "4.9E-324", not "5.0E-324"); it differs from d.0 only for the smallest subnormals
        auto r2 = std::to_chars(buf, buf + sizeof buf, std::fabs(d), std::chars_format::scientific, 1);
        std::string s2(buf, r2.ptr);
        const size_t e2 = s2.find('e');
        dg.clear();
        for (char c : s2.substr(0, e2)) if (c != '.') dg += c;
        if (dg.size() == 2 && dg[1] == '0') dg.resize(1);
        const int exp2 = std::stoi(s2.substr(e2 + 1));
        if (exp2 != exp10) { dg = std::string(1, digits[0]); }   // rounding carried into the next decade: keep d.0
    }
    if (neg) out += '-';
    const double a = std::fabs(d);
    if (a >= 1e-3 && a < 1e7) {
        if (exp10 >= 0) {
            const size_t ip = static_cast<size_t>(exp10) + 1;           // digits of the integer part
            if (dg.size() <= ip) {
                out += dg;
                out.append(ip - dg.size(), '0');
                out += ".0";
            } else {
                out.append(dg, 0, ip);
                out += '.';
                out.append(dg, ip, std::string::npos);
            }
        } else {
            out += "0.";
            out.append(static_cast<size_t>(-exp10 - 1), '0');
            out += dg;
        }
    } else {
        out += dg[0];
        out += '.';
        if (dg.size() > 1) out.append(dg, 1, std::string::npos);
        else out += '0';
        out += 'E';
        out += std::to_string(exp10);
    }
}

struct NewickArgs {
    int32_t n;
    const int32_t *left, *right, *parent;
    const double* height;
    const int32_t* nstubs;
    const double* rate;
    int32_t n_meta;
    std::vector<std::string> ids;
    const double* meta;
    const int32_t* meta_dim;
};

// STL:77-199, statement by statement (the reversible-jump stub-location block STL:103-123 prints only the count here)
void to_newick(const NewickArgs& a, int32_t node, std::string& buf) {
    bool first = true;
    if (a.left[node] >= 0) {
        buf += '(';
        to_newick(a, a.left[node], buf);
        if (a.right[node] >= 0) {
            buf += ',';
            to_newick(a, a.right[node], buf);
        }
        buf += ')';
    } else {
        buf += std::to_string(node + 1);                 // STL:87
    }
    std::string b2 = "[&";
    const bool root = a.parent[node] < 0;
    if (a.nstubs && !root) {                             // STL:92-133
        b2 += "nstubs=";
        b2 += std::to_string(a.nstubs[node]);
        first = false;
    }
    if (a.rate) {                                        // STL:136-141
        if (!first) b2 += ',';
        b2 += "rate=";
        java_double(a.rate[node], b2);
        first = false;
    }
    if (a.n_meta > 0 && !first) b2 += ',';               // STL:144-146
    for (int32_t k = 0; k < a.n_meta; k++) {             // STL:147-187
        if (a.meta_dim[k] > node) {
            b2 += a.ids[k];
            b2 += '=';
            java_double(a.meta[static_cast<int64_t>(k) * a.n + node], b2);
        }
        if (b2.size() > 2 && k < a.n_meta - 1) b2 += ',';   // STL:184-186
    }
    b2 += ']';
    if (b2.size() > 3) buf += b2;                        // STL:191-193
    buf += ':';
    java_double(root ? 0.0 : a.height[a.parent[node]] - a.height[node], buf);   // Node.getLength
}

// ---- the inverse: one logged tree back into arrays (what util/GetEffectiveRates.java reads through BEAST's tree-log parser) ----
struct ParsedNode {
    int32_t left = -1, right = -1, parent = -1, label = -1;
    double length = 0.0;
    std::vector<std::pair<std::string, double>> meta;
};
struct NewickParser {
    const char* p;
    std::vector<ParsedNode> nodes;
    bool ok = true;
    void skip() { while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') ++p; }
    double number() {
        skip();
        // Double.toString spellings incl. NaN / Infinity (strtod accepts "nan" / "inf" / "infinity" case-insensitively)
        char* e = nullptr;
        const double v = strtod(p, &e);
        if (e == p) { ok = false; return 0.0; }
        p = e;
        return v;
    }
    // subtree := ( '(' subtree (',' subtree)* ')' | label ) [ '[&' key=value,... ']' ] [ ':' length ]
    int32_t subtree() {
        skip();
        const int32_t me = static_cast<int32_t>(nodes.size());
        nodes.emplace_back();
        if (*p == '(') {
            ++p;
            int nchild = 0;
            for (;;) {
                const int32_t c = subtree();
                if (!ok) return me;
                nodes[c].parent = me;
                if (nchild == 0) nodes[me].left = c;
                else if (nchild == 1) nodes[me].right = c;
                else { ok = false; return me; }              // BEAST trees are binary
                ++nchild;
                skip();
                if (*p == ',') { ++p; continue; }
                if (*p == ')') { ++p; break; }
                ok = false;
                return me;
            }
        } else {
            char* e = nullptr;
            const long v = strtol(p, &e, 10);              // STL:87 prints the node number + 1
            if (e == p || v < 1) { ok = false; return me; }
            nodes[me].label = static_cast<int32_t>(v - 1);
            p = e;
        }
        skip();
        if (p[0] == '[' && p[1] == '&') {
            p += 2;
            for (;;) {
                skip();
                if (*p == ']') { ++p; break; }
                const char* k0 = p;
                while (*p && *p != '=' && *p != ',' && *p != ']') ++p;
                if (*p != '=') { ok = false; return me; }
                std::string key(k0, p);
                ++p;
                const double v = number();
                if (!ok) return me;
                nodes[me].meta.emplace_back(std::move(key), v);
                skip();
                if (*p == ',') ++p;
            }
        }
        skip();
        if (*p == ':') {
            ++p;
            nodes[me].length = number();
        }
        return me;
    }
};

}  // namespace

extern "C" {

int64_t gsm_parse_newick(const char* newick, int32_t cap_nodes, int32_t* parent, int32_t* left, int32_t* right, double* length,
                         double* height, int32_t n_keys, const char* keys, double* values) {
    if (!newick || cap_nodes <= 0 || !parent || n_keys < 0 || (n_keys > 0 && (!keys || !values))) return -1;
    NewickParser ps{newick};
    const int32_t root = ps.subtree();
    ps.skip();
    if (!ps.ok || (*ps.p != 0 && *ps.p != ';')) return -1;
    const int32_t n = static_cast<int32_t>(ps.nodes.size());
    if (n > cap_nodes || (n & 1) == 0) return n > cap_nodes ? n : -1;
    const int32_t L = (n + 1) / 2;
    // BEAST numbering: a leaf takes the number its label gives it, internal nodes follow in post-order, the root last
    std::vector<int32_t> nr(n, -1);
    std::vector<char> seen(L, 0);
    int32_t next = L;
    // post-order without recursion: parse order is pre-order, children have larger parse indices than their parent
    std::vector<int32_t> order;
    order.reserve(n);
    {
        std::vector<std::pair<int32_t, int>> st;
        st.emplace_back(root, 0);
        while (!st.empty()) {
            auto& [v, state] = st.back();
            const ParsedNode& nd = ps.nodes[v];
            if (state == 0) { state = 1; if (nd.left >= 0) { st.emplace_back(nd.left, 0); continue; } }
            if (state == 1) { state = 2; if (nd.right >= 0) { st.emplace_back(nd.right, 0); continue; } }
            order.push_back(v);
            st.pop_back();
        }
    }
    for (int32_t v : order) {
        const ParsedNode& nd = ps.nodes[v];
        if (nd.left < 0) {
            if (nd.label < 0 || nd.label >= L || seen[nd.label]) return -1;
            seen[nd.label] = 1;
            nr[v] = nd.label;
        } else {
            if (nd.right < 0 || next >= n) return -1;
            nr[v] = next++;
        }
    }
    std::vector<std::string> ks;
    const char* q = keys;
    for (int32_t k = 0; k < n_keys; k++) {
        const char* e = q;
        while (*e && *e != '\n') e++;
        ks.emplace_back(q, e);
        q = *e ? e + 1 : e;
    }
    const double nan = std::nan("");
    for (int32_t k = 0; k < n_keys; k++)
        for (int32_t i = 0; i < n; i++) values[static_cast<int64_t>(k) * cap_nodes + i] = nan;
    // depth of every node below the root, then heights with the youngest leaf at 0
    std::vector<double> depth(n, 0.0);
    double deepest = 0.0;
    for (auto it = order.rbegin(); it != order.rend(); ++it) {   // reverse post-order: parents before children
        const int32_t v = *it;
        const ParsedNode& nd = ps.nodes[v];
        depth[v] = nd.parent < 0 ? 0.0 : depth[nd.parent] + nd.length;
        if (nd.left < 0 && depth[v] > deepest) deepest = depth[v];
    }
    for (int32_t v = 0; v < n; v++) {
        const ParsedNode& nd = ps.nodes[v];
        const int32_t i = nr[v];
        parent[i] = nd.parent < 0 ? -1 : nr[nd.parent];
        if (left) left[i] = nd.left < 0 ? -1 : nr[nd.left];
        if (right) right[i] = nd.right < 0 ? -1 : nr[nd.right];
        if (length) length[i] = nd.parent < 0 ? 0.0 : nd.length;
        if (height) height[i] = deepest - depth[v];
        for (const auto& kv : nd.meta)
            for (int32_t k = 0; k < n_keys; k++)
                if (kv.first == ks[k]) values[static_cast<int64_t>(k) * cap_nodes + i] = kv.second;
    }
    return n;
}

int gsm_effective_leaf_rates(int32_t n_leaves, const double* rate, const double* spike, const double* length, double* out) {
    if (n_leaves < 0 || !rate || !spike || !length || !out) return GSM_E_ARG;
    for (int32_t i = 0; i < n_leaves; i++) out[i] = rate[i] + spike[i] / length[i];   // GetEffectiveRates.java:58-61
    return 0;
}

int64_t gsm_java_double_to_string(double d, char* out, int64_t cap) {
    std::string s;
    java_double(d, s);
    if (out && cap > 0) {
        const size_t w = std::min(static_cast<size_t>(cap - 1), s.size());
        memcpy(out, s.data(), w);
        out[w] = 0;
    }
    return static_cast<int64_t>(s.size());
}

int64_t gsm_format_newick(int32_t n_nodes, const int32_t* left, const int32_t* right, const int32_t* parent, const double* height,
                          const int32_t* nstubs, const double* rate, int32_t n_meta, const char* meta_ids, const double* meta_values,
                          const int32_t* meta_dim, char* out, int64_t cap) {
    if (n_nodes <= 0 || !left || !right || !parent || !height || n_meta < 0 || (n_meta > 0 && (!meta_ids || !meta_values || !meta_dim)))
        return -1;
    NewickArgs a{n_nodes, left, right, parent, height, nstubs, rate, n_meta, {}, meta_values, meta_dim};
    const char* p = meta_ids;
    for (int32_t k = 0; k < n_meta; k++) {               // ids separated by '\n'
        const char* q = p;
        while (*q && *q != '\n') q++;
        a.ids.emplace_back(p, q);
        p = *q ? q + 1 : q;
    }
    int32_t root = -1;
    for (int32_t i = 0; i < n_nodes; i++) {
        if (parent[i] < 0) root = i;
        if (left[i] >= n_nodes || right[i] >= n_nodes || parent[i] >= n_nodes) return -1;
    }
    if (root < 0) return -1;
    std::string s;
    s.reserve(static_cast<size_t>(n_nodes) * 96);
    to_newick(a, root, s);
    if (out && cap > 0) {
        const size_t w = std::min(static_cast<size_t>(cap - 1), s.size());
        memcpy(out, s.data(), w);
        out[w] = 0;
    }
    return static_cast<int64_t>(s.size());
}

}  // extern "C"
