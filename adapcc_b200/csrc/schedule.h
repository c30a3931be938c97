// Strategy model: lenient XML reader, tree roles, relay control, per-rank tree plans.
//
// Parity targets in the reference:
//   * XML -> roles (treeDFS / getStrategyFromXML): /root/reference/csrc/allreduce.cu:52-104
//   * relay controller truth table:                /root/reference/csrc/control.cu:7-101
//   * tinyxml2 (vendored, 5.4 kLoC) is replaced by a ~150-line reader that accepts the
//     same malformed attribute lists the reference ships (strategy/4.xml:3 has
//     id='1'ip='...' with no separating space).
// Fixed on purpose: slot order mismatches cannot occur (children are
// addressed by rank, parents pull), tails are never dropped, active children only.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.h"

namespace adapcc {

struct XmlNode {
  std::string name;
  std::vector<std::pair<std::string, std::string>> attrs;
  std::vector<XmlNode> children;
  const std::string* attr(const std::string& k) const {
    for (auto& kv : attrs)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
};

// Returns false (and sets the error string) on unrecoverable syntax errors.
bool parse_xml(const std::string& text, XmlNode* root);

struct StrategyTree {
  int root = -1;
  std::vector<int> nodes;                 // DFS pre-order (document order)
  std::map<int, int> parent;              // child rank -> parent rank (root absent)
  std::map<int, std::vector<int>> children;  // document order
  std::map<int, std::string> ip;
};

struct Strategy {
  std::vector<StrategyTree> trees;
  // Parse <trees><root id ip><gpu id ip>...; ranks >= world (if world > 0) are contracted
  // out of every tree (their children are re-attached to the nearest remaining ancestor).
  bool load(const std::string& xml_text, int world);
  bool load_file(const std::string& path, int world);
};

// Reference-compatible relay decision for one rank in one tree.
struct RelayControl {
  bool has_recv = false, has_local = false, has_kernel = false, has_send = false;
  std::vector<int> active_recvs;          // children whose subtree contains an active rank
};
RelayControl relay_control(const StrategyTree& tree, int rank, const std::vector<bool>& active);

enum RelayMode : int { RELAY_FORWARD = 0, RELAY_BYPASS = 1 };

// Host-side mirror of the kernel's TreeRole (kept POD-free of CUDA headers for tests).
struct HostTreeRole {
  int parent = -1;
  std::vector<int> children;
  int flags = 0;                          // TreeRoleFlags
  bool any() const { return flags != 0; }
};

// Per-rank role in `tree` for primitive `prim` given the active set. In RELAY_BYPASS mode
// inactive ranks are contracted out first (uniform NVSwitch: nobody needs a forwarder).
HostTreeRole tree_role(const StrategyTree& tree, int rank, const std::vector<bool>& active,
                       int prim, int relay_mode);

}  // namespace adapcc
