// Host-side (CPU) preparation of the read-only tables the kernels consume ("K0" in DESIGN.md):
// Newick ingest with the reference's numbering / time rules, epoch slicing, LCA-time matrix.
//
// Reference behaviour mirrored (file:line under src/main/java/uc/sgp/sagephy/):
//   io/NewickStringAlgorithms.java:27-48, io/NewickTreeReader.java:99-109,158-296   Newick grammar, HOST= capture
//   io/NewickVertex.java:367-375                                                     post-order numbering
//   io/PrIMENewickTreeVerifier.java:237-270                                          vertex times from branch lengths
//   topology/RBTreeEpochDiscretiser.java:146-221,225-229                             epochs, LCA distance
//   apps/SaGePhy/GuestTreeInHostTreeCreator.java:76-86                               stem override, 0 -> 1e-64
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace sgp {

struct SgpError : std::runtime_error { using std::runtime_error::runtime_error; };

// Flat parsed Newick tree; vertex ids are the reference's post-order numbers.
struct FlatTree {
    int n = 0, root = -1;
    std::vector<int> parent, first_child, next_sibling, n_children;
    std::vector<std::string> name; std::vector<uint8_t> has_name;
    std::vector<double> bl;        std::vector<uint8_t> has_bl;
    std::vector<std::string> meta; std::vector<uint8_t> has_meta;
    std::vector<int> host_tag;     // -10 when absent
    std::string tree_meta; bool has_tree_meta = false;
    std::vector<int> bfs;          // NewickTree.getVerticesAsList order
    std::vector<int> children_of(int v) const { std::vector<int> c; for (int x = first_child[v]; x >= 0; x = next_sibling[x]) c.push_back(x); return c; }
    bool is_leaf(int v) const { return n_children[v] == 0; }
};

FlatTree parse_newick(const std::string& text);

// Everything a sampler needs to know about one host (species or gene) tree.
struct HostModel {
    int nV = 0, nE = 0, root = -1, nLeaves = 0;
    std::string tree_name; bool has_tree_name = false;
    std::vector<int> parent, left, right, host_tag, v2e;
    std::vector<double> vt_raw, at_raw;     // TimesMap vertex / arc times (printed in the .guest2host header)
    std::vector<double> vtime;              // discretised vertex time = lower time of the epoch above the vertex
    std::vector<double> ep_lo, ep_mid, ep_up;
    std::vector<int> ep_off, ep_arcs;       // arcs of epoch e: ep_arcs[ep_off[e] .. ep_off[e+1])
    std::vector<std::string> leaf_name;     // "" for interior vertices
    std::vector<int> leaves;                // RBTree.getLeaves(): ids in increasing order
    std::vector<double> lca_time;           // nV*nV, only filled when requested
    int max_arcs = 0;
    std::string header;                     // RBTreeEpochDiscretiser.toString()

    // stem: nullptr = keep the value in the tree; otherwise override (GuestTreeGen always passes one, default 0).
    void build(const FlatTree& t, const std::string* tree_name, const double* stem, bool need_lca);
    bool is_root(int x) const { return parent[x] < 0; }
    bool is_leaf(int x) const { return left[x] < 0; }
};

}  // namespace sgp
