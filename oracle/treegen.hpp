// ORACLE — test infrastructure only (see oracle/README.md). PARITY UNPINNED.
//
// Guest-tree-in-host-tree sampler, replicate driver, pruning, vertex tags and the text reports. Restates
//   apps/SaGePhy/GuestVertex.java:11-33,89-119,183-425           (events, node record, setMeta)
//   apps/SaGePhy/GuestTreeInHostTreeCreator.java:110-321,332-363,372-423,431-559  (createUnprunedTree, findVertex,
//        isAncestor, createGuestVertex, getInfo, getLeafMap, getSigma)
//   apps/SaGePhy/PruningHelper.java:25-145                        (labelling, prune)
//   apps/SaGePhy/GuestTreeMachina.java:69-169                     (attempt loop, unprunedIsOK)
//   apps/SaGePhy/HostTreeGen.java:124-174                         (HostTreeGen's own getInfo)
//   math/ExponentialDistribution.java:79-81,155-159
#pragma once
#include <deque>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "hostmodel.hpp"

namespace oracle {

enum class Event : uint8_t {
    SPECIATION, CODIVERGENCE, LEAF, UNSAMPLED_LEAF, DUPLICATION, LOSS, REPLACING_LOSS, ADDITIVE_TRANSFER, REPLACING_TRANSFER,
    AT_INTRAGENE_INTRASPECIES, AT_INTRAGENE_INTERSPECIES, AT_INTERGENE_INTRASPECIES, AT_INTERGENE_INTERSPECIES,
    RT_INTRAGENE_INTRASPECIES, RT_INTRAGENE_INTERSPECIES, RT_INTERGENE_INTRASPECIES, RT_INTERGENE_INTERSPECIES
};
inline const char* event_java_name(Event e) {      // Event.toString() == enum constant name (used by getSigma)
    static const char* n[] = {"SPECIATION", "CODIVERGENCE", "LEAF", "UNSAMPLED_LEAF", "DUPLICATION", "LOSS", "REPLACING_LOSS",
        "ADDITIVE_TRANSFER", "REPLACING_TRANSFER", "ADDITIVE_TRANSFER_INTRAGENE_INTRASPECIES", "ADDITIVE_TRANSFER_INTRAGENE_INTERSPECIES",
        "ADDITIVE_TRANSFER_INTERGENE_INTRASPECIES", "ADDITIVE_TRANSFER_INTERGENE_INTERSPECIES", "REPLACING_TRANSFER_INTRAGENE_INTRASPECIES",
        "REPLACING_TRANSFER_INTRAGENE_INTERSPECIES", "REPLACING_TRANSFER_INTERGENE_INTRASPECIES", "REPLACING_TRANSFER_INTERGENE_INTERSPECIES"};
    return n[(int)e];
}
enum class Prun : uint8_t { UNPRUNABLE, PRUNABLE, COLLAPSABLE, UNKNOWN };

struct GVertex {
    // NewickVertex part
    int number = -1;
    const std::string* name_prefix = nullptr; bool name_sigma = false;   // name = prefix + number [+ "_" + sigma], built on demand
    double bl = 0;
    std::string meta;             // "" (present but empty) until setMeta
    GVertex* parent = nullptr;
    GVertex* lc = nullptr; GVertex* rc = nullptr; int nchildren = 0;   // children == null <=> nchildren == 0
    // GuestVertex part
    Event event = Event::LEAF;
    double abstime = 0;
    int sigma = -1;
    int fromArc = -1, toArc = -1;
    const std::string* fromGuest = nullptr; const std::string* toGuest = nullptr;     // null -> printed "null"
    HostModel* tree = nullptr;    // guestTree (the host model this vertex lives in)
    EpochO* epoch = nullptr;
    Prun prun = Prun::UNKNOWN;
    bool is_leaf() const { return nchildren == 0; }
    GVertex* left() const { return nchildren == 0 ? nullptr : lc; }
    GVertex* right() const { return nchildren <= 1 ? nullptr : rc; }
    void set_children(GVertex* a, GVertex* b) { lc = a; rc = b; nchildren = 2; }
    void clear_children() { lc = rc = nullptr; nchildren = 0; }
    std::string name() const {
        if (!name_prefix) return std::string();
        std::string s = *name_prefix + std::to_string(number);
        if (name_sigma) { s.push_back('_'); s += std::to_string(sigma); }
        return s;
    }
};

// Vertex pool reused across attempts (the reference allocates fresh objects; pooling only makes the CPU baseline faster).
struct Arena {
    std::deque<GVertex> pool; size_t used = 0;
    GVertex* make() {
        if (used == pool.size()) pool.emplace_back();
        GVertex* v = &pool[used++];
        v->number = -1; v->name_prefix = nullptr; v->name_sigma = false; v->meta.clear(); v->parent = v->lc = v->rc = nullptr; v->nchildren = 0;
        v->fromArc = v->toArc = -1; v->fromGuest = v->toGuest = nullptr; v->prun = Prun::UNKNOWN;
        return v;
    }
    void clear() { used = 0; }
};

struct MaxAttemptsException : std::runtime_error { using std::runtime_error::runtime_error; };

// interface UnprunedGuestTreeCreator (apps/SaGePhy/UnprunedGuestTreeCreator.java:17-59)
class Creator {
public:
    virtual ~Creator() {}
    virtual GVertex* create_unpruned_tree(JavaMT& prng, Arena& arena) = 0;
    virtual std::vector<int> host_leaves() const = 0;
    virtual std::string info(const GVertex* root, bool doML) const = 0;
    virtual std::string leaf_map(const GVertex* root) const = 0;
    virtual std::string sigma(const GVertex* root) const = 0;
    virtual bool any_unused() const { return false; }      // getUsed() contains a zero (DomainTreeGen -all)
};

inline double exp_sample(JavaMT& prng, double lambda) { double x = prng.nextDouble(); return -j_log(1.0 - x) / lambda; }

class GuestInHostCreator : public Creator {
public:
    HostModel& host;
    double lambda, mu, tau, theta; std::string bias; double bias_rate; bool doGeneBirth; double gene_birth; bool ishost; double rho;
    uint64_t vertices_created = 0;

    GuestInHostCreator(HostModel& h, double lambda_, double mu_, double tau_, double theta_, std::string bias_, double bias_rate_,
                       bool gb, double gbc, bool ishost_, double rho_)
        : host(h), lambda(lambda_), mu(mu_), tau(tau_), theta(theta_), bias(std::move(bias_)), bias_rate(bias_rate_), doGeneBirth(gb),
          gene_birth(gbc), ishost(ishost_), rho(rho_) {
        if (lambda < 0 || mu < 0 || tau < 0) throw std::invalid_argument("IllegalArgumentException: Cannot have rate less than 0.");
        if (rho < 0 || rho > 1) throw std::invalid_argument("IllegalArgumentException: Cannot have leaf sampling probability outside [0,1].");
    }

    GVertex* create_guest_vertex(int X, double startTime, JavaMT& prng, Arena& arena) {      // :372-423
        // Epoch.sampleArc returns -5 when no arc of the epoch carries the donor's HOST tag (topology/Epoch.java:256-258,273-275);
        // the reference then indexes its host arrays with -5 and dies with an ArrayIndexOutOfBoundsException
        if (X < 0) throw std::out_of_range("ArrayIndexOutOfBoundsException: " + std::to_string(X));
        ++vertices_created;
        bool isRoot = host.is_root(X);
        double sum = isRoot ? lambda + mu : lambda + mu + tau;
        if (sum == 0.0) sum = 1e-48;
        double lowerTime = host.vertex_time(X);
        double branchTime = exp_sample(prng, sum);
        double eventTime = startTime - branchTime;
        Event event; EpochO* epoch;
        if (eventTime <= lowerTime) {
            eventTime = lowerTime;
            branchTime = round_sig(startTime - eventTime, 8);
            if (host.is_leaf(X)) event = (prng.nextDouble() < rho ? Event::LEAF : Event::UNSAMPLED_LEAF);
            else event = Event::SPECIATION;
            epoch = &host.epochs[host.epoch_no_above(X)];
        } else {
            double rnd = prng.nextDouble();
            if (isRoot) {
                event = (rnd < lambda / sum) ? Event::DUPLICATION : Event::LOSS;
            } else {
                if (rnd >= (mu + tau) / sum) event = Event::DUPLICATION;
                else if (rnd < mu / sum) event = Event::LOSS;
                else { double trn = prng.nextDouble(); event = (trn < theta) ? Event::REPLACING_TRANSFER : Event::ADDITIVE_TRANSFER; }
            }
            int epno = host.epoch_no_above(X);
            while (host.epochs[epno].upper() < eventTime) ++epno;
            epoch = &host.epochs[epno];
        }
        GVertex* v = arena.make();
        v->event = event; v->sigma = X; v->epoch = epoch; v->abstime = eventTime; v->bl = branchTime; v->tree = &host;
        return v;
    }

    static GVertex* find_vertex(GVertex* node, const GVertex* x) {       // :332-346
        if (!node) return nullptr;
        if (node->sigma == x->toArc && node != x->right() && node != x->left() && node->abstime < x->abstime && (node->abstime + node->bl) > x->abstime)
            return node;
        GVertex* n = find_vertex(node->left(), x);
        if (!n) n = find_vertex(node->right(), x);
        return n;
    }
    static bool is_ancestor(const GVertex* replaced, const GVertex* node) {      // :349-363
        if (!replaced) return false;
        if (replaced == node) return true;
        return is_ancestor(replaced->left(), node) || is_ancestor(replaced->right(), node);
    }

    GVertex* create_unpruned_tree(JavaMT& prng, Arena& arena) override {     // :110-321
        std::deque<GVertex*> alive;
        std::map<double, GVertex*> sorted;      // TreeMap<Double,GuestVertex>: equal keys overwrite
        GVertex* root; int myRoot;
        if (ishost) { myRoot = host.root; root = create_guest_vertex(myRoot, host.tip_to_leaf_time(), prng, arena); }
        else if (doGeneBirth) { myRoot = host.sample_root(prng, gene_birth); root = create_guest_vertex(myRoot, host.vertex_upper_time(myRoot), prng, arena); }
        else { myRoot = host.root; root = create_guest_vertex(myRoot, host.tip_to_leaf_time(), prng, arena); }
        alive.push_back(root);
        auto release = [&]() { if (alive.empty() && !sorted.empty()) { auto it = std::prev(sorted.end()); alive.push_back(it->second); sorted.erase(it); } };
        while (!alive.empty()) {
            GVertex* lin = alive.front(); alive.pop_front();
            if (lin->event == Event::LOSS || lin->event == Event::REPLACING_LOSS || lin->event == Event::LEAF || lin->event == Event::UNSAMPLED_LEAF) {
                release();
                continue;
            }
            GVertex *lc = nullptr, *rc = nullptr;
            if (lin->event == Event::SPECIATION) {
                lc = create_guest_vertex(host.left[lin->sigma], lin->abstime, prng, arena);
                rc = create_guest_vertex(host.right[lin->sigma], lin->abstime, prng, arena);
            } else if (lin->event == Event::DUPLICATION) {
                lc = create_guest_vertex(lin->sigma, lin->abstime, prng, arena);
                rc = create_guest_vertex(lin->sigma, lin->abstime, prng, arena);
            } else if (lin->event == Event::ADDITIVE_TRANSFER) {
                bool stayLeft = prng.nextDouble() < 0.5;
                GVertex* stay = create_guest_vertex(lin->sigma, lin->abstime, prng, arena);
                int to = host.sample_arc(*lin->epoch, prng, lin->sigma, lin->epoch->find_index_of_arc(lin->sigma), lin->abstime, bias, bias_rate, nullptr, host.host_tag[lin->sigma], true);
                lin->fromArc = lin->sigma;
                GVertex* moved = create_guest_vertex(to, lin->abstime, prng, arena);
                lin->toArc = lin->epoch->transferedToArc;
                if (stayLeft) { lc = stay; rc = moved; } else { rc = stay; lc = moved; }
            } else if (lin->event == Event::REPLACING_TRANSFER) {
                bool stayLeft = prng.nextDouble() < 0.5;
                GVertex* stay = create_guest_vertex(lin->sigma, lin->abstime, prng, arena);
                int hostArc = host.host_tag[lin->sigma];
                int firstTo = host.sample_arc(*lin->epoch, prng, lin->sigma, lin->epoch->find_index_of_arc(lin->sigma), lin->abstime, bias, bias_rate, nullptr, hostArc, true);
                lin->fromArc = lin->sigma;
                lin->toArc = lin->epoch->transferedToArc;
                GVertex* node = find_vertex(root, lin);
                std::vector<int> emptyArcs;
                while (!node && (int)emptyArcs.size() != lin->epoch->n_arcs() - 1) {
                    if (std::find(emptyArcs.begin(), emptyArcs.end(), lin->toArc) == emptyArcs.end()) emptyArcs.push_back(lin->toArc);
                    // a second -5 (no candidate left, list still short of n_arcs - 1) changes nothing any more: the reference spins here forever
                    else if (lin->toArc == -5) throw std::runtime_error("no recipient arc with the donor's HOST tag is left (the reference does not terminate here)");
                    if ((int)emptyArcs.size() != lin->epoch->n_arcs() - 1) {
                        lin->toArc = host.sample_arc(*lin->epoch, prng, lin->sigma, lin->epoch->find_index_of_arc(lin->sigma), lin->abstime, bias, bias_rate, &emptyArcs, hostArc, true);
                        node = find_vertex(root, lin);
                    }
                }
                GVertex* moved;
                if (node) {
                    moved = create_guest_vertex(lin->toArc, lin->abstime, prng, arena);
                    node->event = Event::REPLACING_LOSS;
                    node->abstime = lin->abstime;
                    for (auto it = sorted.begin(); it != sorted.end();) { if (is_ancestor(node, it->second)) it = sorted.erase(it); else ++it; }
                    node->clear_children();
                } else {
                    lin->event = Event::ADDITIVE_TRANSFER;
                    moved = create_guest_vertex(firstTo, lin->abstime, prng, arena);
                    lin->toArc = firstTo;
                    lin->epoch->transferedToArc = firstTo;
                }
                if (stayLeft) { lc = stay; rc = moved; } else { rc = stay; lc = moved; }
            } else {
                throw std::runtime_error("UnsupportedOperationException: Unexpected event type.");
            }
            lin->set_children(lc, rc);
            lc->parent = lin; rc->parent = lin;
            if (lc->event != Event::REPLACING_TRANSFER) alive.push_back(lc); else sorted[lc->abstime] = lc;
            if (rc->event != Event::REPLACING_TRANSFER) alive.push_back(rc); else sorted[rc->abstime] = rc;
            release();
        }
        if (root->bl <= 1.0e-32) root->bl = 0.0;
        return root;
    }

    std::vector<int> host_leaves() const override { return host.leaves(); }

    std::string info(const GVertex* guestRoot, bool doML) const override {      // :431-511
        int nV = 0, nLeaves = 0, nSpecs = 0, nDups = 0, nLoss = 0, nRLoss = 0, nAT = 0, nRT = 0;
        double totalTime = 0.0, beneath = 0.0;
        std::deque<const GVertex*> q; if (guestRoot) q.push_back(guestRoot);
        double hostRootTime = host.vertex_time(host.root);
        while (!q.empty()) {
            const GVertex* v = q.front(); q.pop_front();
            if (!v->is_leaf()) { q.push_back(v->left()); q.push_back(v->right()); }
            ++nV; totalTime += v->bl;
            beneath += std::max(std::min(hostRootTime - v->abstime, v->bl), 0.0);
            switch (v->event) {
                case Event::DUPLICATION: ++nDups; break; case Event::LOSS: ++nLoss; break; case Event::REPLACING_LOSS: ++nRLoss; break;
                case Event::ADDITIVE_TRANSFER: ++nAT; break; case Event::REPLACING_TRANSFER: ++nRT; break; case Event::SPECIATION: ++nSpecs; break;
                case Event::LEAF: case Event::UNSAMPLED_LEAF: ++nLeaves; break;
                default: throw std::runtime_error("UnsupportedOperationException: Unexpected event type.");
            }
        }
        totalTime = round_sig(totalTime, 8); beneath = round_sig(beneath, 8);
        std::string sb;
        auto I = [&](const char* k, int v) { sb += k; sb += std::to_string(v); sb.push_back('\n'); };
        auto D = [&](const char* k, double v) { sb += k; sb += java_double_to_string(v); sb.push_back('\n'); };
        I("No. of vertices:\t", nV); I("No. of extant leaves:\t", nLeaves); I("No. of speciations:\t", nSpecs); I("No. of duplications:\t", nDups);
        I("No. of losses:\t", nLoss); I("No. of replacing losses:\t", nRLoss); I("No. of additive transfers:\t", nAT); I("No. of replacing transfers:\t", nRT);
        D("Total branch time:\t", totalTime); D("Total branch time beneath host stem:\t", beneath);
        if (doML) {
            D("Duplication ML estimate:\t", round_sig(nDups / totalTime, 8)); D("Loss ML estimate:\t", round_sig(nLoss / totalTime, 8));
            D("Replacing Loss ML estimate:\t", round_sig(nRLoss / totalTime, 8)); D("Additive Transfer ML estimate:\t", round_sig(nAT / totalTime, 8));
            D("Replacing Transfer ML estimate:\t", round_sig(nRT / totalTime, 8));
        }
        return sb;
    }

    std::string leaf_map(const GVertex* guestRoot) const override {     // :514-532
        std::string sb; std::deque<const GVertex*> q; if (guestRoot) q.push_back(guestRoot);
        while (!q.empty()) {
            const GVertex* v = q.front(); q.pop_front();
            if (!v->is_leaf()) { q.push_back(v->left()); q.push_back(v->right()); }
            else if (v->event == Event::LEAF || v->event == Event::UNSAMPLED_LEAF) {
                sb += v->name(); sb.push_back('\t'); sb += (host.names[v->sigma] ? *host.names[v->sigma] : std::string("null")); sb.push_back('\n');
            }
        }
        return sb;
    }

    std::string sigma(const GVertex* guestRoot) const override {        // :535-559
        std::string sb = "# GUEST-TO-HOST MAP\n";
        sb += "Host tree:\t" + host.to_string() + "\n";
        sb += "Guest vertex name:\tGuest vertex ID:\tGuest vertex type:\tGuest vertex time:\tHost vertex/arc ID:\tHost epoch ID:\n";
        std::deque<const GVertex*> q; if (guestRoot) q.push_back(guestRoot);
        while (!q.empty()) {
            const GVertex* v = q.front(); q.pop_front();
            if (!v->is_leaf()) { q.push_back(v->left()); q.push_back(v->right()); }
            sb += v->name(); sb.push_back('\t'); sb += std::to_string(v->number); sb.push_back('\t'); sb += event_java_name(v->event); sb.push_back('\t');
            sb += java_double_to_string(v->abstime); sb.push_back('\t'); sb += std::to_string(v->sigma); sb.push_back('\t'); sb += std::to_string(v->epoch->no); sb.push_back('\n');
        }
        return sb;
    }
};

// ---- DomainTreeInHostTreeCreator (apps/SaGePhy/DomainTreeInHostTreeCreator.java:101-1300) ---------------------------------------
// A domain tree evolving inside a forest of gene trees. The eight transfer kinds share one code path here, parameterised by
// (additive | replacing) x (intra | inter gene) x (intra | inter species); the reference spells out all sixteen branches.
class DomainCreator : public Creator {
public:
    HostModel& species; std::vector<HostModel*> genes; std::vector<std::string> gene_names;
    double lambda, mu, tau, theta; std::string bias; double bias_rate; bool doBirth; double birth; double inter_gene, inter_species, rho;
    std::vector<int> used; GVertex* root = nullptr; uint64_t vertices_created = 0;

    DomainCreator(HostModel& sp, std::vector<HostModel*> g, double l, double m, double t, double th, std::string b, double br, bool db, double dbc,
                  double ig, double is, double rho_)
        : species(sp), genes(std::move(g)), lambda(l), mu(m), tau(t), theta(th), bias(std::move(b)), bias_rate(br), doBirth(db), birth(dbc),
          inter_gene(ig), inter_species(is), rho(rho_) {
        if (lambda < 0 || mu < 0 || tau < 0 || birth < 0) throw std::invalid_argument("IllegalArgumentException: Cannot have rate less than 0.");
        if (theta < 0 || theta > 1 || inter_gene < 0 || inter_gene > 1 || inter_species < 0 || inter_species > 1 || rho < 0 || rho > 1)
            throw std::invalid_argument("IllegalArgumentException: Cannot have probability outside [0,1].");
        for (HostModel* h : genes) gene_names.push_back(h->tree_name ? *h->tree_name : std::string("null"));
    }
    int tree_index(const HostModel* h) const { for (size_t i = 0; i < genes.size(); ++i) if (genes[i] == h) return (int)i; return -1; }

    GVertex* create_guest_vertex(int X, double startTime, JavaMT& prng, Arena& arena, HostModel* g) {      // :1042-1120
        ++vertices_created;
        bool isRoot = g->is_root(X);
        double sum = isRoot ? lambda + mu : lambda + mu + tau;
        if (sum == 0.0) sum = 1e-48;
        double lowerTime = g->vertex_time(X);
        double branchTime = exp_sample(prng, sum);
        double eventTime = startTime - branchTime;
        Event event; EpochO* epoch;
        if (eventTime <= lowerTime) {
            eventTime = lowerTime;
            branchTime = round_sig(startTime - eventTime, 8);
            if (g->is_leaf(X)) event = (prng.nextDouble() < rho ? Event::LEAF : Event::UNSAMPLED_LEAF); else event = Event::CODIVERGENCE;
            epoch = &g->epochs[g->epoch_no_above(X)];
        } else {
            double rnd = prng.nextDouble();
            if (isRoot) event = (rnd < lambda / sum) ? Event::DUPLICATION : Event::LOSS;
            else if (rnd >= (mu + tau) / sum) event = Event::DUPLICATION;
            else if (rnd < mu / sum) event = Event::LOSS;
            else {
                double trn = prng.nextDouble(), trn_gen = prng.nextDouble(), trn_spe = prng.nextDouble();
                int base = (trn < theta) ? (int)Event::RT_INTRAGENE_INTRASPECIES : (int)Event::AT_INTRAGENE_INTRASPECIES;
                event = (Event)(base + (trn_gen < inter_gene ? 2 : 0) + (trn_spe < inter_species ? 1 : 0));
            }
            int epno = g->epoch_no_above(X);
            while (g->epochs[epno].upper() < eventTime) ++epno;
            epoch = &g->epochs[epno];
        }
        used[tree_index(g)]++;
        GVertex* v = arena.make();
        v->event = event; v->sigma = X; v->epoch = epoch; v->abstime = eventTime; v->bl = branchTime; v->tree = g;
        return v;
    }

    static GVertex* find_vertex(GVertex* node, const GVertex* x, const HostModel* g) {     // :948-964
        if (!node) return nullptr;
        if (node->sigma == x->toArc && node->tree == g && node != x->right() && node != x->left() && node->abstime < x->abstime && (node->abstime + node->bl) > x->abstime) return node;
        GVertex* n = find_vertex(node->left(), x, g);
        if (!n) n = find_vertex(node->right(), x, g);
        return n;
    }

    // sampleInterGeneArc :1004-1033. Returns (-5, null) when there is no candidate.
    std::pair<int, HostModel*> sample_inter_gene_arc(GVertex* lin, bool intraSpecies, const std::vector<int>* emptyArcs, JavaMT& prng) {
        std::vector<std::pair<int, HostModel*>> cand;
        int hostArc = lin->tree->host_tag[lin->sigma];
        for (HostModel* tr : genes) {
            if (tr == lin->tree) continue;
            for (int i = 0; i < tr->k; ++i) {
                int arcHost = tr->host_tag[i];
                if (i != lin->sigma && (!emptyArcs || std::find(emptyArcs->begin(), emptyArcs->end(), i) == emptyArcs->end()) &&
                    tr->vertex_time(i) < lin->abstime && tr->vertex_upper_time(i) > lin->abstime && (intraSpecies ? (arcHost == hostArc) : (arcHost != hostArc)))
                    cand.push_back({i, tr});
            }
        }
        if (cand.empty()) return {-5, nullptr};
        int c = prng.nextInt((int)cand.size());
        lin->epoch->transferedToArc = cand[c].first;
        return cand[c];
    }

    GVertex* create_unpruned_tree(JavaMT& prng, Arena& arena) override {      // :158-937
        std::deque<GVertex*> alive; std::map<double, GVertex*> sorted;
        int first_guest = prng.nextInt((int)genes.size());
        used.assign(genes.size(), 0);
        HostModel* g0 = genes[first_guest];
        if (doBirth) { int r0 = g0->sample_root(prng, birth); root = create_guest_vertex(r0, g0->vertex_upper_time(r0), prng, arena, g0); }
        else root = create_guest_vertex(g0->root, g0->tip_to_leaf_time(), prng, arena, g0);
        alive.push_back(root);
        auto release = [&]() { if (alive.empty() && !sorted.empty()) { auto it = std::prev(sorted.end()); alive.push_back(it->second); sorted.erase(it); } };
        auto is_repl = [](Event e) { return (int)e >= (int)Event::RT_INTRAGENE_INTRASPECIES; };
        while (!alive.empty()) {
            GVertex* lin = alive.front(); alive.pop_front();
            if (lin->event == Event::LOSS || lin->event == Event::REPLACING_LOSS || lin->event == Event::LEAF || lin->event == Event::UNSAMPLED_LEAF) { release(); continue; }
            GVertex *lc = nullptr, *rc = nullptr;
            if (lin->event == Event::CODIVERGENCE) {
                lc = create_guest_vertex(lin->tree->left[lin->sigma], lin->abstime, prng, arena, lin->tree);
                rc = create_guest_vertex(lin->tree->right[lin->sigma], lin->abstime, prng, arena, lin->tree);
            } else if (lin->event == Event::DUPLICATION) {
                lc = create_guest_vertex(lin->sigma, lin->abstime, prng, arena, lin->tree);
                rc = create_guest_vertex(lin->sigma, lin->abstime, prng, arena, lin->tree);
            } else if ((int)lin->event >= (int)Event::AT_INTRAGENE_INTRASPECIES) {
                const int code = ((int)lin->event - (int)Event::AT_INTRAGENE_INTRASPECIES) & 3;
                const bool replacing = is_repl(lin->event), inter = (code & 2) != 0, include = (code & 1) == 0;
                const bool stayLeft = prng.nextDouble() < 0.5;
                const int hostArc = lin->tree->host_tag[lin->sigma];
                int to; HostModel* totree;
                if (!inter) { to = lin->tree->sample_arc(*lin->epoch, prng, lin->sigma, 0, lin->abstime, bias, bias_rate, nullptr, hostArc, include); totree = lin->tree; }
                else { auto pr = sample_inter_gene_arc(lin, include, nullptr, prng); to = pr.first; totree = pr.second; }
                const int firstTo = to; HostModel* origGene = totree;
                if (to == -5) {
                    // swap (:984-1002): the transfer vertex is replaced in place by a freshly sampled vertex
                    GVertex* child = create_guest_vertex(lin->sigma, lin->abstime, prng, arena, lin->tree);
                    if (lin != root) { GVertex* par = lin->parent; if (par->lc == lin) par->lc = child; else par->rc = child; } else root = child;
                    if (child != root) child->parent = lin->parent;
                    lin->parent = nullptr; lin = nullptr;
                    if (stayLeft) lc = child; else rc = child;
                } else {
                    GVertex* stay = create_guest_vertex(lin->sigma, lin->abstime, prng, arena, lin->tree);
                    GVertex* moved;
                    lin->fromArc = lin->sigma; lin->toArc = lin->epoch->transferedToArc;
                    if (!replacing) moved = create_guest_vertex(to, lin->abstime, prng, arena, totree);
                    else {
                        GVertex* node = find_vertex(root, lin, totree);
                        std::vector<int> emptyArcs;
                        while (!node && (int)emptyArcs.size() != lin->epoch->n_arcs() - 1 && lin->toArc != -5) {
                            if (std::find(emptyArcs.begin(), emptyArcs.end(), lin->toArc) == emptyArcs.end()) emptyArcs.push_back(lin->toArc);
                            if ((int)emptyArcs.size() != lin->epoch->n_arcs() - 1) {
                                if (!inter) lin->toArc = lin->tree->sample_arc(*lin->epoch, prng, lin->sigma, 0, lin->abstime, bias, bias_rate, &emptyArcs, hostArc, include);
                                else { auto pr = sample_inter_gene_arc(lin, include, &emptyArcs, prng); lin->toArc = pr.first; totree = pr.second; }
                                node = find_vertex(root, lin, totree);
                            }
                        }
                        if (node) {
                            moved = create_guest_vertex(lin->toArc, lin->abstime, prng, arena, totree);
                            node->event = Event::REPLACING_LOSS; node->abstime = lin->abstime;
                            for (auto it = sorted.begin(); it != sorted.end();) { if (GuestInHostCreator::is_ancestor(node, it->second)) it = sorted.erase(it); else ++it; }
                            node->clear_children();
                        } else {
                            moved = create_guest_vertex(firstTo, lin->abstime, prng, arena, origGene);
                            lin->event = (Event)((int)lin->event - 4);      // the additive kind with the same gene / species scope
                            lin->toArc = firstTo; lin->epoch->transferedToArc = firstTo;
                        }
                    }
                    lin->fromGuest = &gene_names[tree_index(lin->tree)]; lin->toGuest = &gene_names[tree_index(moved->tree)];
                    if (stayLeft) { lc = stay; rc = moved; } else { rc = stay; lc = moved; }
                }
            } else throw std::runtime_error("UnsupportedOperationException: Unexpected event type.");
            if (lin) { lin->set_children(lc, rc); lc->parent = lin; rc->parent = lin; }
            if (lc) { if (!is_repl(lc->event)) alive.push_back(lc); else sorted[lc->abstime] = lc; }
            if (rc) { if (!is_repl(rc->event)) alive.push_back(rc); else sorted[rc->abstime] = rc; }
            release();
        }
        if (root->bl <= 1.0e-32) root->bl = 0.0;
        return root;
    }

    std::vector<int> host_leaves() const override { return species.leaves(); }
    bool any_unused() const override { for (int u : used) if (u == 0) return true; return false; }

    std::string info(const GVertex* guestRoot, bool doML) const override {       // :1128-1252
        int cnt[17] = {0}; int nV = 0; double totalTime = 0.0, beneath = 0.0;
        std::deque<const GVertex*> q; if (guestRoot) q.push_back(guestRoot);
        double hostRootTime = species.vertex_time(species.root);
        while (!q.empty()) {
            const GVertex* v = q.front(); q.pop_front();
            if (!v->is_leaf()) { q.push_back(v->left()); q.push_back(v->right()); }
            ++nV; totalTime += v->bl; beneath += std::max(std::min(hostRootTime - v->abstime, v->bl), 0.0);
            if (v->event == Event::SPECIATION || v->event == Event::ADDITIVE_TRANSFER || v->event == Event::REPLACING_TRANSFER) throw std::runtime_error("UnsupportedOperationException: Unexpected event type.");
            cnt[(int)v->event]++;
        }
        totalTime = round_sig(totalTime, 8); beneath = round_sig(beneath, 8);
        std::string sb;
        auto I = [&](const std::string& k, int v) { sb += k; sb += std::to_string(v); sb.push_back('\n'); };
        auto D = [&](const std::string& k, double v) { sb += k; sb += java_double_to_string(v); sb.push_back('\n'); };
        static const char* scope[4] = {"intragene intraspecies", "intragene interspecies", "intergene intraspecies", "intergene interspecies"};
        static const char* Scope[4] = {"Intragene Intraspecies", "Intragene Interspecies", "Intergene Intraspecies", "Intergene Interspecies"};
        I("No. of vertices:\t", nV); I("No. of extant leaves:\t", cnt[(int)Event::LEAF] + cnt[(int)Event::UNSAMPLED_LEAF]);
        I("No. of co-divergences:\t", cnt[(int)Event::CODIVERGENCE]); I("No. of duplications:\t", cnt[(int)Event::DUPLICATION]);
        I("No. of losses:\t", cnt[(int)Event::LOSS]); I("No. of replacing losses:\t", cnt[(int)Event::REPLACING_LOSS]);
        for (int k = 0; k < 4; ++k) I(std::string("No. of additive transfers ") + scope[k] + ":\t", cnt[(int)Event::AT_INTRAGENE_INTRASPECIES + k]);
        for (int k = 0; k < 4; ++k) I(std::string("No. of replacing transfers ") + scope[k] + ":\t", cnt[(int)Event::RT_INTRAGENE_INTRASPECIES + k]);
        D("Total branch time:\t", totalTime); D("Total branch time beneath host stem:\t", beneath);
        if (doML) {
            D("Duplication ML estimate:\t", round_sig(cnt[(int)Event::DUPLICATION] / totalTime, 8)); D("Loss ML estimate:\t", round_sig(cnt[(int)Event::LOSS] / totalTime, 8));
            D("Replacing Loss ML estimate:\t", round_sig(cnt[(int)Event::REPLACING_LOSS] / totalTime, 8));
            for (int k = 0; k < 4; ++k) D(std::string("Additive Transfer ") + Scope[k] + " ML estimate:\t", round_sig(cnt[(int)Event::AT_INTRAGENE_INTRASPECIES + k] / totalTime, 8));
            for (int k = 0; k < 4; ++k) D(std::string("Replacing Transfer ") + Scope[k] + " ML estimate:\t", round_sig(cnt[(int)Event::RT_INTRAGENE_INTRASPECIES + k] / totalTime, 8));
        }
        return sb;
    }
    std::string leaf_map(const GVertex* guestRoot) const override {      // :1255-1273 (three columns)
        std::string sb; std::deque<const GVertex*> q; if (guestRoot) q.push_back(guestRoot);
        while (!q.empty()) {
            const GVertex* v = q.front(); q.pop_front();
            if (!v->is_leaf()) { q.push_back(v->left()); q.push_back(v->right()); }
            else if (v->event == Event::LEAF || v->event == Event::UNSAMPLED_LEAF) {
                sb += v->name(); sb.push_back('\t'); sb += (v->tree->names[v->sigma] ? *v->tree->names[v->sigma] : std::string("null")); sb.push_back('\t');
                sb += (v->tree->tree_name ? *v->tree->tree_name : std::string("null")); sb.push_back('\n');
            }
        }
        return sb;
    }
    std::string sigma(const GVertex* guestRoot) const override {         // :1276-1300 (host tree = gene tree of the root vertex)
        std::string sb = "# GUEST-TO-HOST MAP\n";
        sb += "Host tree:\t" + guestRoot->tree->to_string() + "\n";
        sb += "Guest vertex name:\tGuest vertex ID:\tGuest vertex type:\tGuest vertex time:\tHost vertex/arc ID:\tHost epoch ID:\n";
        std::deque<const GVertex*> q; q.push_back(guestRoot);
        while (!q.empty()) {
            const GVertex* v = q.front(); q.pop_front();
            if (!v->is_leaf()) { q.push_back(v->left()); q.push_back(v->right()); }
            sb += v->name(); sb.push_back('\t'); sb += std::to_string(v->number); sb.push_back('\t'); sb += event_java_name(v->event); sb.push_back('\t');
            sb += java_double_to_string(v->abstime); sb.push_back('\t'); sb += std::to_string(v->sigma); sb.push_back('\t'); sb += std::to_string(v->epoch->no); sb.push_back('\n');
        }
        return sb;
    }
};

// ---- PruningHelper ------------------------------------------------------------------------------------------------------
inline void set_label(GVertex* v, int no, const std::string& prefix, bool appendSigma) { v->number = no; v->name_prefix = &prefix; v->name_sigma = appendSigma; }
inline int label_unprunable(GVertex* v, int nextNo, const std::string& prefix, bool appendSigma) {      // PruningHelper.java:25-65
    if (v->event == Event::LOSS || v->event == Event::REPLACING_LOSS || v->event == Event::UNSAMPLED_LEAF) v->prun = Prun::PRUNABLE;
    else if (v->event == Event::LEAF) { v->prun = Prun::UNPRUNABLE; set_label(v, nextNo++, prefix, appendSigma); }
    else {
        // (a single-child vertex cannot occur on this path: children are always set in pairs)
        GVertex *lc = v->lc, *rc = v->rc;
        nextNo = label_unprunable(lc, nextNo, prefix, appendSigma);
        nextNo = label_unprunable(rc, nextNo, prefix, appendSigma);
        if (lc->prun == Prun::PRUNABLE && rc->prun == Prun::PRUNABLE) v->prun = Prun::PRUNABLE;
        else if (lc->prun == Prun::PRUNABLE || rc->prun == Prun::PRUNABLE) v->prun = Prun::COLLAPSABLE;
        else { v->prun = Prun::UNPRUNABLE; set_label(v, nextNo++, prefix, appendSigma); }
    }
    return nextNo;
}
inline int label_prunable(GVertex* v, int nextNo, const std::string& prefix, bool appendSigma) {        // :76-90
    if (!v->is_leaf()) { nextNo = label_prunable(v->lc, nextNo, prefix, appendSigma); nextNo = label_prunable(v->rc, nextNo, prefix, appendSigma); }
    if (v->number == -1) set_label(v, nextNo++, prefix, appendSigma);
    return nextNo;
}
inline GVertex* prune(const GVertex* u, Arena& arena) {      // :98-145
    if (u->prun == Prun::PRUNABLE) return nullptr;
    auto copy = [&](const GVertex* o) { GVertex* c = arena.make(); *c = *o; return c; };   // shallow copy incl. children refs
    if (u->event == Event::LEAF) return copy(u);
    GVertex* lc = prune(u->left(), arena);
    GVertex* rc = prune(u->right(), arena);
    if (!lc) { rc->bl = round_sig(rc->bl + u->bl, 8); return rc; }
    if (!rc) { lc->bl = round_sig(lc->bl + u->bl, 8); return lc; }
    GVertex* v = copy(u);
    v->set_children(lc, rc); lc->parent = v; rc->parent = v;
    return v;
}

// ---- GuestVertex.setMeta (apps/SaGePhy/GuestVertex.java:183-425) -------------------------------------------------------------
inline void set_meta(GVertex* root, bool isSpecies) {
    std::deque<GVertex*> q; q.push_back(root);
    while (!q.empty()) {
        GVertex* v = q.front(); q.pop_front();
        std::string sb = "[ID=" + std::to_string(v->number);
        if (!isSpecies) { sb += " HOST=" + std::to_string(v->sigma); sb += " GUEST=" + (v->tree->tree_name ? *v->tree->tree_name : std::string("null")); }
        auto discpt = [&]() { int j = 0; while (j < 3) { if (v->epoch->times[j] >= v->abstime) break; ++j; } return "DISCPT=(" + std::to_string(v->epoch->no) + "," + std::to_string(j) + ")"; };
        auto fromto2 = [&]() { return "(" + std::to_string(v->fromArc) + "," + std::to_string(v->toArc) + ")"; };
        auto fromto4 = [&]() { return "(" + std::to_string(v->fromArc) + "," + std::to_string(v->toArc) + ";" + (v->fromGuest ? *v->fromGuest : std::string("null")) + "," + (v->toGuest ? *v->toGuest : std::string("null")) + ")"; };
        switch (v->event) {
            case Event::DUPLICATION: sb += std::string(isSpecies ? " VERTEXTYPE=Speciation " : " VERTEXTYPE=Duplication ") + discpt(); break;
            case Event::LOSS: sb += " VERTEXTYPE=Loss"; break;
            case Event::REPLACING_LOSS: sb += " VERTEXTYPE=Replacing Loss"; break;
            case Event::ADDITIVE_TRANSFER: sb += " VERTEXTYPE=Additive Transfer FROMTOLINEAGE=" + fromto2() + " " + discpt(); break;
            case Event::REPLACING_TRANSFER: sb += " VERTEXTYPE=Replacing Transfer FROMTOLINEAGE=" + fromto2() + " " + discpt(); break;
            case Event::AT_INTRAGENE_INTRASPECIES: sb += " VERTEXTYPE=Additive Transfer Intragene Intraspecies FROMTOLINEAGE=" + fromto4() + " " + discpt(); break;
            case Event::AT_INTRAGENE_INTERSPECIES: sb += " VERTEXTYPE=Additive Transfer Intragene Interspecies FROMTOLINEAGE=" + fromto4() + " " + discpt(); break;
            case Event::AT_INTERGENE_INTRASPECIES: sb += " VERTEXTYPE=Additive Transfer Intergene Intraspecies FROMTOLINEAGE=" + fromto4() + " " + discpt(); break;
            case Event::AT_INTERGENE_INTERSPECIES: sb += " VERTEXTYPE=Additive Transfer Intergene Interspecies FROMTOLINEAGE=" + fromto4() + " " + discpt(); break;
            case Event::RT_INTRAGENE_INTRASPECIES: sb += " VERTEXTYPE=Replacing Transfer Intragene Intraspecies FROMTOLINEAGE=" + fromto4() + " " + discpt(); break;
            case Event::RT_INTRAGENE_INTERSPECIES: sb += " VERTEXTYPE=Replacing Transfer Intragene Interspecies FROMTOLINEAGE=" + fromto4() + " " + discpt(); break;
            case Event::RT_INTERGENE_INTRASPECIES: sb += " VERTEXTYPE=Replacing Transfer Intergene Intraspecies FROMTOLINEAGE=" + fromto4() + " " + discpt(); break;
            case Event::RT_INTERGENE_INTERSPECIES: sb += " VERTEXTYPE=Replacing Transfer Intergene Interspecies FROMTOLINEAGE=" + fromto4() + " " + discpt(); break;
            case Event::SPECIATION: sb += " VERTEXTYPE=Speciation " + discpt(); break;
            case Event::CODIVERGENCE: sb += " VERTEXTYPE=Co-divergence " + discpt(); break;
            case Event::LEAF: sb += " VERTEXTYPE=Leaf"; break;
            case Event::UNSAMPLED_LEAF: sb += " VERTEXTYPE=UnsampledLeaf"; break;
        }
        sb += "]";
        v->meta = sb;
        if (!v->is_leaf()) { q.push_back(v->lc); q.push_back(v->rc); }
    }
}

inline void gv_to_string(const GVertex* v, std::string& sb) {    // NewickVertex.toString with name/bl/meta always present
    if (v->nchildren > 0) { sb.push_back('('); gv_to_string(v->lc, sb); sb.push_back(','); gv_to_string(v->rc, sb); sb.push_back(')'); }
    sb += v->name(); sb.push_back(':'); sb += java_double_to_string(v->bl); sb += v->meta;
}
inline std::string gtree_to_string(const GVertex* root, const char* treeMeta) {
    std::string s; gv_to_string(root, s); if (treeMeta) s += treeMeta; s.push_back(';'); return s;
}

// ---- GuestTreeMachina ------------------------------------------------------------------------------------------------------
struct MachinaResult { GVertex* pruned = nullptr; GVertex* unpruned = nullptr; int attempts = 0; };

class Machina {
public:
    JavaMT prng; int min, max, minper, maxper; const std::vector<int>* leafSizes; int maxAttempts; std::string prefix;
    bool excludeMeta, appendSigma, isSpecies; int attempts = 0;
    Arena arena;      // vertices of the current attempt (+ pruned copies)
    std::vector<const GVertex*> ok_queue;

    Machina(const std::string& seed, int min_, int max_, int minper_, int maxper_, const std::vector<int>* sizes, int maxAttempts_,
            std::string prefix_, bool excludeMeta_, bool appendSigma_, bool isSpecies_)
        : prng(seed), min(min_), max(max_), minper(minper_), maxper(maxper_), leafSizes(sizes), maxAttempts(maxAttempts_),
          prefix(std::move(prefix_)), excludeMeta(excludeMeta_), appendSigma(appendSigma_), isSpecies(isSpecies_) {}

    bool unpruned_is_ok(const GVertex* root, int exact, const std::vector<int>& hostLeaves, const Creator& creator, bool all_genes) {     // :124-169
        int sampled = 0; std::map<int, int> cnt;
        ok_queue.clear(); if (root) ok_queue.push_back(root);
        for (size_t qi = 0; qi < ok_queue.size(); ++qi) {
            const GVertex* v = ok_queue[qi];
            if (v->event == Event::LEAF) { ++sampled; ++cnt[v->sigma]; }
            else if (!v->is_leaf()) { ok_queue.push_back(v->lc); ok_queue.push_back(v->rc); }
        }
        if (exact != -1 && sampled != exact) return false;
        if (sampled < min || sampled > max) return false;
        for (int l : hostLeaves) { auto it = cnt.find(l); int c = it == cnt.end() ? 0 : it->second; if (c < minper || c > maxper) return false; }
        if (all_genes && creator.any_unused()) return false;
        return true;
    }

    MachinaResult sample(Creator& creator, bool all_genes) {        // sampleGuestTree :69-115
        attempts = 0;
        std::vector<int> hostLeaves = creator.host_leaves();
        int exact = -1;
        if (leafSizes) exact = (*leafSizes)[prng.nextInt((int)leafSizes->size())];
        GVertex* unpruned;
        do {
            if (attempts > maxAttempts) throw MaxAttemptsException(std::to_string(attempts) + " reached.");
            arena.clear();
            unpruned = creator.create_unpruned_tree(prng, arena);
            int no = label_unprunable(unpruned, 0, prefix, appendSigma);
            label_prunable(unpruned, no, prefix, appendSigma);
            ++attempts;
        } while (!unpruned_is_ok(unpruned, exact, hostLeaves, creator, all_genes));
        if (!excludeMeta) set_meta(unpruned, isSpecies);
        MachinaResult r; r.unpruned = unpruned; r.pruned = prune(unpruned, arena);
        ++attempts;
        r.attempts = attempts;
        return r;
    }
};

// HostTreeGen.getInfo (apps/SaGePhy/HostTreeGen.java:124-174)
inline std::string host_tree_gen_info(const GVertex* root, bool doML) {
    int nV = 0, nB = 0, nD = 0, nL = 0; double totalTime = 0.0;
    std::deque<const GVertex*> q; if (root) q.push_back(root);
    while (!q.empty()) {
        const GVertex* v = q.front(); q.pop_front();
        if (!v->is_leaf()) { q.push_back(v->left()); q.push_back(v->right()); }
        ++nV; totalTime += v->bl;
        switch (v->event) {
            case Event::DUPLICATION: case Event::SPECIATION: ++nB; break;
            case Event::LOSS: ++nD; break;
            case Event::LEAF: case Event::UNSAMPLED_LEAF: ++nL; break;
            default: break;
        }
    }
    totalTime = round_sig(totalTime, 8);
    std::string sb;
    sb += "No. of vertices:\t" + std::to_string(nV) + "\n";
    sb += "No. of extant leaves:\t" + std::to_string(nL) + "\n";
    sb += "No. of births:\t" + std::to_string(nB) + "\n";
    sb += "No. of deaths:\t" + std::to_string(nD) + "\n";
    sb += "Total branch time:\t" + java_double_to_string(totalTime) + "\n";
    if (doML) {
        sb += "Birth ML estimate:\t" + java_double_to_string(round_sig(nB / totalTime, 8)) + "\n";
        sb += "Death ML estimate:\t" + java_double_to_string(round_sig(nD / totalTime, 8)) + "\n";
    }
    return sb;
}

}  // namespace oracle
