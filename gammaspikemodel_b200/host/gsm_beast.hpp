// gsm_beast.hpp — C++ host-side mirror of the reference's BEAST 2 plugin classes for the hot path, written
// above the C ABI of include/gsm_b200.h.  The reference is Java (BEAST 2.7.7); this image has no JVM, so this
// header plays the role the modified Java classes play in a BEAST install (INTEGRATION.md shows the Panama
// binding): same class names, same method names, same argument meaning and the same error behaviour
// (invalid state -> -infinity from calculateLogP, std::invalid_argument / std::runtime_error where the Java
// throws IllegalArgumentException / RuntimeException at initAndValidate).  Every number comes from the device:
// nothing here evaluates a branch term on the CPU.
//
// One Model = one chain state (a batch of one tree); the batched entry points are the gsm_* functions themselves.
//
// Reference citations: PRCM = src/gammaspike/clockmodel/PunctuatedRelaxedClockModel.java, BSP / BRP / STP =
// src/gammaspike/distribution/{BranchSpikePrior,BranchRatePrior,StumpedTreePrior}.java, Stubs =
// src/gammaspike/tree/Stubs.java, SPL = src/gammaspike/logger/SaltativeProportionLogger.java.
#pragma once

#include <cctype>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "gsm_b200.h"

namespace gammaspike {

// ---- beast.base.evolution.tree.Tree / TreeParser (array form) -------------------------------------------------
struct Tree {
    std::vector<double> height;
    std::vector<int32_t> parent;   // -1 = root
    int leafCount = 0;

    int getNodeCount() const { return static_cast<int>(height.size()); }
    int getLeafNodeCount() const { return leafCount; }
    int getRootNr() const {
        for (int i = 0; i < getNodeCount(); i++) if (parent[i] < 0) return i;
        return -1;
    }
    double getRootHeight() const { return height[getRootNr()]; }

    // TreeParser(newick): IsLabelledNewick, adjustTipHeights = false, offset = 1 -> taxon "k" is node k-1
    // (SURVEY.md Appendix A).  Leaves first, internal nodes in post-order, root last.
    static Tree parseNewick(const std::string& s) {
        struct N { int left = -1, right = -1, par = -1; double len = 0; std::string label; };
        std::vector<N> nodes;
        size_t pos = 0;
        auto skip = [&] { while (pos < s.size() && std::isspace(static_cast<unsigned char>(s[pos]))) pos++; };
        std::function<int()> parse = [&]() -> int {
            skip();
            int id = static_cast<int>(nodes.size());
            nodes.emplace_back();
            if (pos < s.size() && s[pos] == '(') {
                pos++;
                int l = parse();
                skip();
                if (s[pos] != ',') throw std::invalid_argument("newick: expected ','");
                pos++;
                int r = parse();
                skip();
                if (s[pos] != ')') throw std::invalid_argument("newick: expected ')'");
                pos++;
                nodes[id].left = l; nodes[id].right = r; nodes[l].par = id; nodes[r].par = id;
            }
            skip();
            size_t st = pos;
            while (pos < s.size() && s[pos] != ':' && s[pos] != ',' && s[pos] != ')' && s[pos] != ';' && !std::isspace(static_cast<unsigned char>(s[pos]))) pos++;
            nodes[id].label = s.substr(st, pos - st);
            skip();
            if (pos < s.size() && s[pos] == ':') {
                pos++;
                skip();
                size_t used = 0;
                nodes[id].len = std::stod(s.substr(pos), &used);
                pos += used;
            }
            return id;
        };
        int root = parse();
        int n = static_cast<int>(nodes.size());
        std::vector<double> depth(n, 0.0);
        std::vector<int> order;   // pre-order
        std::vector<int> stack{root};
        while (!stack.empty()) {
            int v = stack.back(); stack.pop_back(); order.push_back(v);
            if (nodes[v].left >= 0) {
                depth[nodes[v].left] = depth[v] + nodes[nodes[v].left].len;
                depth[nodes[v].right] = depth[v] + nodes[nodes[v].right].len;
                stack.push_back(nodes[v].right); stack.push_back(nodes[v].left);
            }
        }
        double maxd = 0;
        for (double d : depth) maxd = std::max(maxd, d);
        std::vector<int> nr(n, -1);
        int leaves = 0;
        for (int v = 0; v < n; v++) if (nodes[v].left < 0) leaves++;
        int seq = 0;
        for (int v : order) if (nodes[v].left < 0) {
            bool numeric = !nodes[v].label.empty();
            for (char c : nodes[v].label) numeric = numeric && std::isdigit(static_cast<unsigned char>(c));
            nr[v] = numeric ? std::stoi(nodes[v].label) - 1 : seq;
            seq++;
        }
        int next = leaves;
        std::function<void(int)> post = [&](int v) {
            if (nodes[v].left < 0) return;
            post(nodes[v].left); post(nodes[v].right);
            nr[v] = next++;
        };
        post(root);
        Tree t;
        t.leafCount = leaves;
        t.height.assign(n, 0.0);
        t.parent.assign(n, -1);
        for (int v = 0; v < n; v++) {
            t.height[nr[v]] = maxd - depth[v];
            t.parent[nr[v]] = nodes[v].par < 0 ? -1 : nr[nodes[v].par];
        }
        return t;
    }
};

// ---- the device context shared by the plugin objects of one chain -------------------------------------------
class Model {
public:
    explicit Model(const Tree& tree, int device = 0) : tree_(tree) {
        const int n = tree_.getNodeCount();
        if (gsm_create(&h_, device, 1, n) != GSM_OK) throw std::runtime_error(std::string("gsm_create: ") + gsm_last_error(nullptr));
        rates.assign(n, 1.0);
        spikes.assign(n, 1.0);
        nstubs.assign(n, 0);
        params.assign(GSM_NPARAM, 0.0);
        params[GSM_P_CLOCK_RATE] = 1.0; params[GSM_P_SPIKE_MEAN] = 0.01; params[GSM_P_SPIKE_SHAPE] = 2.0;
        params[GSM_P_CLOCK_SD] = 0.5; params[GSM_P_LAMBDA] = 2.0; params[GSM_P_MU] = 1.0; params[GSM_P_RHO] = 1.0;
    }
    ~Model() { gsm_destroy(h_); }
    Model(const Model&) = delete;
    Model& operator=(const Model&) = delete;

    Tree& tree() { return tree_; }
    std::vector<double> rates, spikes, params;
    std::vector<int32_t> nstubs;
    uint8_t useSpike = 1;
    uint32_t flags = 0;
    int32_t stubMode = GSM_STUBS_COUNTS;

    // evaluate the current state on the device; result row GSM_O_*
    const double* eval(bool want_rates = false) {
        const int n = tree_.getNodeCount();
        check(gsm_set_mode(h_, stubMode, flags));
        check(gsm_upload_trees(h_, 1, n, tree_.height.data(), tree_.parent.data(), nullptr));
        check(gsm_upload_branch_params(h_, rates.data(), spikes.data(), nstubs.data()));
        check(gsm_upload_tree_params(h_, params.data(), &useSpike));
        if (want_rates) { out_rates.resize(n); out_wspikes.resize(n); }
        check(gsm_eval(h_, out, want_rates ? out_rates.data() : nullptr, want_rates ? out_wspikes.data() : nullptr));
        return out;
    }
    std::vector<double> cdf(int nodeNr) {
        eval();
        std::vector<double> buf(4096);
        int32_t len = 0;
        check(gsm_stub_cdf(h_, 0, nodeNr, buf.data(), static_cast<int32_t>(buf.size()), &len));
        buf.resize(std::min<int32_t>(len, 4096));
        return buf;
    }
    double out[GSM_NOUT];
    std::vector<double> out_rates, out_wspikes;

private:
    void check(int rc) { if (rc != GSM_OK) throw std::runtime_error(std::string("gsm_b200: ") + gsm_last_error(h_)); }
    Tree tree_;
    gsm_handle* h_ = nullptr;
};

// ---- gammaspike.tree.Stubs (mode selection Stubs:138-245; count accessors Stubs:344-370) ---------------------
class Stubs {
public:
    explicit Stubs(Model& m) : m_(m) {}
    void setStubsPerBranch(const std::vector<int32_t>& v) {              // branchCounts mode, dimension N-1 (Stubs:200)
        m_.stubMode = GSM_STUBS_COUNTS;
        for (size_t i = 0; i < m_.nstubs.size(); i++) m_.nstubs[i] = i < v.size() ? v[i] : 0;
    }
    void setPrior() { m_.stubMode = GSM_STUBS_NOSTUB; }                  // nostub mode: integrate over counts (Stubs:143-189)
    bool estimateStubs() const { return m_.stubMode != GSM_STUBS_NOSTUB; }   // Stubs:648-650
    int getNStubsOnBranch(int nodeNr) const {                            // Stubs:344-370
        if (m_.stubMode == GSM_STUBS_NOSTUB) return -1;
        if (m_.stubMode == GSM_STUBS_COUNTS && nodeNr >= static_cast<int>(m_.nstubs.size()) - 1) return 0;
        return m_.nstubs[nodeNr];
    }
private:
    Model& m_;
};

// ---- gammaspike.clockmodel.PunctuatedRelaxedClockModel ---------------------------------------------------------
class PunctuatedRelaxedClockModel {
public:
    explicit PunctuatedRelaxedClockModel(Model& m) : m_(m) {}
    void setIndicator(bool v) { m_.flags |= GSM_F_HAS_INDICATOR; m_.useSpike = v ? 1 : 0; }    // PRCM:34
    void setRates(const std::vector<double>& r) { m_.flags |= GSM_F_HAS_RATES; m_.rates = r; }  // PRCM:36
    void setSpikes(const std::vector<double>& s) { m_.spikes = s; }                             // PRCM:37
    void setSpikeMean(double v) { m_.params[GSM_P_SPIKE_MEAN] = v; }                            // PRCM:33
    void setClockRate(double v) { m_.params[GSM_P_CLOCK_RATE] = v; }                            // clock.rate
    void setRelaxed(bool v) { if (v) m_.flags &= ~GSM_F_RELAXED_FALSE; else m_.flags |= GSM_F_RELAXED_FALSE; }   // PRCM:35
    void setNoSpikeOnDatedTips(bool v) { if (v) m_.flags |= GSM_F_NO_SPIKE_DATED_TIPS; else m_.flags &= ~GSM_F_NO_SPIKE_DATED_TIPS; }
    void setStubs(Stubs*) { m_.flags |= GSM_F_STUBS_TO_CLOCK; }                                 // PRCM:30
    void initAndValidate() {                                                                     // PRCM:49-76 (dimension checks)
        const size_t n = m_.tree().getNodeCount();
        if (m_.rates.size() != n || m_.spikes.size() != n) throw std::invalid_argument("rates/spikes must have one entry per node");
    }
    double getRateForBranch(int nodeNr) { m_.eval(true); return m_.out_rates[nodeNr]; }         // PRCM:225-242
    std::vector<double> getRatesArray() { m_.eval(true); return m_.out_rates; }                 // PRCM:267-276
    double getBurstSize(int nodeNr) { m_.eval(true); return m_.out_wspikes[nodeNr]; }           // PRCM:167-208
private:
    Model& m_;
};

// ---- gammaspike.clockmodel.SpikeSize (GSMweightedSpikes) --------------------------------------------------------
class SpikeSize {
public:
    explicit SpikeSize(Model& m) : m_(m) {}
    int getDimension() const { return static_cast<int>(m_.spikes.size()); }
    double getArrayValue(int dim) { m_.eval(true); return m_.out_wspikes[dim]; }                // SpikeSize:48-56
private:
    Model& m_;
};

// ---- gammaspike.distribution.BranchSpikePrior --------------------------------------------------------------------
class BranchSpikePrior {
public:
    explicit BranchSpikePrior(Model& m) : m_(m) {}
    void setShape(double v) { m_.params[GSM_P_SPIKE_SHAPE] = v; }                               // BSP:31
    void setStubs(Stubs*) { m_.flags |= GSM_F_STUBS_TO_SPIKE_PRIOR; }                           // BSP:28
    void setIndicator(bool v) { m_.flags |= GSM_F_BSP_HAS_INDICATOR; m_.useSpike = v ? 1 : 0; } // BSP:35
    double calculateLogP() { return m_.eval()[GSM_O_SPIKE_LOGP]; }                              // BSP:53-187
    // BSP:315-389; mu is StumpedTreePrior.getMeanStubNumber(h_node, h_parent), computed on the device as
    // Stubs.sampleNStubsOnBranch does (Stubs:424-447)
    std::vector<double> getCumulativeProbs(int nodeNr) { return m_.cdf(nodeNr); }
private:
    Model& m_;
};

// ---- gammaspike.distribution.BranchRatePrior --------------------------------------------------------------------
class BranchRatePrior {
public:
    explicit BranchRatePrior(Model& m) : m_(m) {}
    void setSigma(double v) { m_.params[GSM_P_CLOCK_SD] = v; }                                  // BRP:23
    double calculateLogP() { m_.flags |= GSM_F_HAS_RATES; return m_.eval()[GSM_O_RATE_LOGP]; }  // BRP:29-65
private:
    Model& m_;
};

// ---- gammaspike.distribution.StumpedTreePrior ------------------------------------------------------------------
class StumpedTreePrior {
public:
    explicit StumpedTreePrior(Model& m) : m_(m) {}
    // BEASTObject.setInputValue(name, value) for the Inputs of STP:24-50; unknown names throw like the Java does
    void setInputValue(const std::string& name, double v) {
        static const char* known[] = {"origin", "lambda", "r0", "psi", "netDiversificationRate", "turnover", "samplingProportion", "rho"};
        for (const char* k : known) if (name == k) { in_[name] = v; return; }
        throw std::invalid_argument("This BEASTInterface (StumpedTreePrior) has no input with name " + name);
    }
    void setInputValue(const std::string& name, bool v) {
        if (name == "conditionOnSampling") condSampling_ = v;
        else if (name == "conditionOnRhoSampling") condRho_ = v;
        else if (name == "conditionOnRoot") condRoot_ = v;
        else if (name == "ignoreTreePrior") ignoreTree_ = v;
        else if (name == "ignoreStubPrior") ignoreStub_ = v;
        else throw std::invalid_argument("This BEASTInterface (StumpedTreePrior) has no input with name " + name);
    }
    void setStubs(Stubs*) { stubs_ = true; }                                                     // STP:46
    bool has(const char* k) const { return in_.count(k) != 0; }

    void initAndValidate() {                                                                     // STP:70-111
        if (has("psi") && has("samplingProportion")) throw std::runtime_error("Please specifiy either psi or samplingProportion or neither, but not both");
        if (has("origin") && condRoot_) throw std::runtime_error("Remove origin or set conditionOnRoot input to \"false\"");
        if (condSampling_ && condRho_) throw std::invalid_argument("Either set to \"true\" only one of conditionOnSampling and conditionOnRhoSampling inputs or don't specify both!");
        if (condRoot_ && !condRho_ && !condSampling_) throw std::invalid_argument("When conditioning on the root, set either conditionOnSampling or conditionOnRhoSampling to true.");
        if (!has("lambda") && !(has("netDiversificationRate") && has("turnover"))) throw std::invalid_argument("lambda or netDiversificationRate+turnover required");
        if (!has("r0") && !has("turnover")) throw std::invalid_argument("r0 or turnover required");
        if (has("origin") && in_["origin"] < m_.tree().getRootHeight()) throw std::invalid_argument("Initial value of origin should be greater than initial root height");
        nExtantTaxa_ = -1;                                                                       // STP:111
    }
    double getLambda() const { return has("lambda") ? in_.at("lambda") : in_.at("netDiversificationRate") / (1 - in_.at("turnover")); }   // STP:672-690
    double getMu() const { return has("r0") ? getLambda() / in_.at("r0") : in_.at("turnover") * getLambda(); }                            // STP:692-708
    double getPsi() const {                                                                                                               // STP:710-728
        if (!has("psi") && !has("samplingProportion")) return 0.0;
        if (!has("psi")) { double s = in_.at("samplingProportion"); return s * getMu() / (1 - s); }
        return in_.at("psi");
    }
    double getRho() const { return has("rho") ? in_.at("rho") : 1.0; }                           // STP:730-737

    double calculateTreeLogLikelihood(const Tree& tree) {                                        // STP:150-284
        m_.tree() = tree;
        m_.params[GSM_P_LAMBDA] = getLambda(); m_.params[GSM_P_MU] = getMu();
        m_.params[GSM_P_PSI] = getPsi(); m_.params[GSM_P_RHO] = getRho();
        m_.params[GSM_P_ORIGIN] = has("origin") ? in_.at("origin") : 0.0;
        uint32_t f = m_.flags & ~(GSM_F_ORIGIN | GSM_F_COND_ROOT | GSM_F_COND_SAMPLING | GSM_F_COND_RHO_SAMPLING |
                                  GSM_F_IGNORE_TREE_PRIOR | GSM_F_IGNORE_STUB_PRIOR | GSM_F_STUBS_TO_TREE_PRIOR);
        if (has("origin")) f |= GSM_F_ORIGIN;
        if (condRoot_) f |= GSM_F_COND_ROOT;
        if (condSampling_) f |= GSM_F_COND_SAMPLING;
        if (condRho_) f |= GSM_F_COND_RHO_SAMPLING;
        if (ignoreTree_) f |= GSM_F_IGNORE_TREE_PRIOR;
        if (ignoreStub_) f |= GSM_F_IGNORE_STUB_PRIOR;
        if (stubs_) f |= GSM_F_STUBS_TO_TREE_PRIOR;
        m_.flags = f;
        const double* o = m_.eval();
        // the cross-call nExtantTaxa guard (STP:196-200) is host state: only on the unconditioned variants
        if (!ignoreTree_ && !has("origin") && !condRoot_ && getLambda() > getMu()) {
            const int n = static_cast<int>(o[GSM_O_N_EXTANT]);
            if (nExtantTaxa_ == -1) nExtantTaxa_ = n;
            else if (n != nExtantTaxa_) return -std::numeric_limits<double>::infinity();
        }
        return o[GSM_O_STP_LOGP];
    }
    double calculateLogP() { return calculateTreeLogLikelihood(m_.tree()); }                     // SpeciesTreeDistribution.calculateLogP
private:
    Model& m_;
    mutable std::map<std::string, double> in_;
    bool condSampling_ = false, condRho_ = false, condRoot_ = false, ignoreTree_ = false, ignoreStub_ = false, stubs_ = false;
    int nExtantTaxa_ = -1;
};

// ---- gammaspike.logger.SaltativeProportionLogger ---------------------------------------------------------------
class SaltativeProportionLogger {
public:
    explicit SaltativeProportionLogger(Model& m) : m_(m) {}
    std::string header(const std::string& clockID) const { return clockID + ".ProportionOfSaltation"; }   // SPL:27-30
    double log() { const double* o = m_.eval(true); return o[GSM_O_SUM_BURST] / o[GSM_O_SUM_DIST]; }      // SPL:33-61
private:
    Model& m_;
};

}  // namespace gammaspike
