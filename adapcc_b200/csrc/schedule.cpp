#include "schedule.h"

#include <algorithm>
#include <fstream>
#include <functional>
#include <set>
#include <sstream>

namespace adapcc {

// ------------------------------------------------------------------------------------
// lenient XML reader
// ------------------------------------------------------------------------------------
namespace {
struct Cursor {
  const std::string& s;
  size_t i = 0;
  explicit Cursor(const std::string& t) : s(t) {}
  bool eof() const { return i >= s.size(); }
  char peek() const { return s[i]; }
  bool starts(const char* lit) const { return s.compare(i, strlen(lit), lit) == 0; }
  void skip_ws() { while (!eof() && isspace((unsigned char)s[i])) ++i; }
};

bool is_name_char(char c) { return isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; }

// skips text, comments, processing instructions and doctype up to the next element tag
bool skip_misc(Cursor& c) {
  while (!c.eof()) {
    if (c.peek() != '<') { ++c.i; continue; }
    if (c.starts("<!--")) {
      size_t e = c.s.find("-->", c.i + 4);
      if (e == std::string::npos) { set_error("xml: unterminated comment"); return false; }
      c.i = e + 3;
    } else if (c.starts("<?")) {
      size_t e = c.s.find("?>", c.i + 2);
      if (e == std::string::npos) { set_error("xml: unterminated <?"); return false; }
      c.i = e + 2;
    } else if (c.starts("<!")) {
      size_t e = c.s.find('>', c.i);
      if (e == std::string::npos) { set_error("xml: unterminated <!"); return false; }
      c.i = e + 1;
    } else {
      return true;
    }
  }
  return true;
}

bool parse_element(Cursor& c, XmlNode* out, int depth) {
  if (depth > 256) { set_error("xml: nesting too deep"); return false; }
  // at '<'
  ++c.i;
  size_t b = c.i;
  while (!c.eof() && is_name_char(c.peek())) ++c.i;
  out->name = c.s.substr(b, c.i - b);
  if (out->name.empty()) { set_error("xml: empty tag name at %zu", b); return false; }
  // attributes; separators between attributes are optional (id='1'ip='x')
  while (true) {
    c.skip_ws();
    if (c.eof()) { set_error("xml: unterminated tag <%s", out->name.c_str()); return false; }
    if (c.starts("/>")) { c.i += 2; return true; }
    if (c.peek() == '>') { ++c.i; break; }
    size_t kb = c.i;
    while (!c.eof() && is_name_char(c.peek())) ++c.i;
    std::string key = c.s.substr(kb, c.i - kb);
    if (key.empty()) { set_error("xml: bad attribute in <%s> at %zu", out->name.c_str(), c.i); return false; }
    c.skip_ws();
    std::string val;
    if (!c.eof() && c.peek() == '=') {
      ++c.i;
      c.skip_ws();
      if (c.eof()) { set_error("xml: dangling '='"); return false; }
      char qc = c.peek();
      if (qc == '"' || qc == '\'') {
        size_t e = c.s.find(qc, c.i + 1);
        if (e == std::string::npos) { set_error("xml: unterminated attribute value"); return false; }
        val = c.s.substr(c.i + 1, e - c.i - 1);
        c.i = e + 1;
      } else {  // unquoted value
        size_t vb = c.i;
        while (!c.eof() && !isspace((unsigned char)c.peek()) && c.peek() != '>' && !c.starts("/>")) ++c.i;
        val = c.s.substr(vb, c.i - vb);
      }
    }
    for (const auto& kv : out->attrs)
      if (kv.first == key) { set_error("xml: attribute '%s' given twice on <%s>", key.c_str(), out->name.c_str()); return false; }
    out->attrs.emplace_back(key, val);
  }
  // children until the matching close tag
  while (true) {
    if (!skip_misc(c)) return false;
    if (c.eof()) { set_error("xml: missing </%s>", out->name.c_str()); return false; }
    if (c.starts("</")) {
      size_t e = c.s.find('>', c.i);
      if (e == std::string::npos) { set_error("xml: unterminated close tag"); return false; }
      c.i = e + 1;
      return true;
    }
    out->children.emplace_back();
    if (!parse_element(c, &out->children.back(), depth + 1)) return false;
  }
}
}  // namespace

bool parse_xml(const std::string& text, XmlNode* root) {
  Cursor c(text);
  if (!skip_misc(c)) return false;
  if (c.eof()) { set_error("xml: no root element"); return false; }
  return parse_element(c, root, 0);
}

// ------------------------------------------------------------------------------------
// strategy trees
// ------------------------------------------------------------------------------------
static bool add_subtree(const XmlNode& x, int parent, StrategyTree* t) {
  const std::string* id = x.attr("id");
  if (!id) { set_error("strategy: <%s> without id", x.name.c_str()); return false; }
  int rank = atoi(id->c_str());
  if (rank < 0) { set_error("strategy: negative rank id"); return false; }
  if (t->parent.count(rank) || rank == t->root) {
    set_error("strategy: rank %d appears twice in one tree", rank);
    return false;
  }
  if (parent < 0) t->root = rank;
  else { t->parent[rank] = parent; t->children[parent].push_back(rank); }
  t->nodes.push_back(rank);
  const std::string* ip = x.attr("ip");
  t->ip[rank] = ip ? *ip : std::string();
  for (const XmlNode& ch : x.children)
    if (ch.name == "gpu" && !add_subtree(ch, rank, t)) return false;
  return true;
}

// keep only ranks satisfying `keep`; re-attach orphans to the nearest kept ancestor
static StrategyTree contract(const StrategyTree& t, const std::function<bool(int)>& keep) {
  StrategyTree o;
  std::vector<int> tops;
  for (int x : t.nodes) {
    if (!keep(x)) continue;
    o.nodes.push_back(x);
    auto ipit = t.ip.find(x);
    o.ip[x] = ipit == t.ip.end() ? std::string() : ipit->second;
    int a = x;
    int found = -1;
    while (true) {
      auto it = t.parent.find(a);
      if (it == t.parent.end()) break;
      a = it->second;
      if (keep(a)) { found = a; break; }
    }
    if (found >= 0) { o.parent[x] = found; o.children[found].push_back(x); }
    else tops.push_back(x);
  }
  if (tops.empty()) return o;
  o.root = keep(t.root) ? t.root : tops[0];
  for (int x : tops)
    if (x != o.root) { o.parent[x] = o.root; o.children[o.root].push_back(x); }
  return o;
}

bool Strategy::load(const std::string& xml_text, int world) {
  trees.clear();
  XmlNode doc;
  if (!parse_xml(xml_text, &doc)) return false;
  if (doc.name != "trees") { set_error("strategy: root element is <%s>, expected <trees>", doc.name.c_str()); return false; }
  for (const XmlNode& r : doc.children) {
    if (r.name != "root") continue;
    StrategyTree t;
    if (!add_subtree(r, -1, &t)) return false;
    if (world > 0) t = contract(t, [world](int x) { return x < world; });
    // The Python front end rejects trees that do not span every rank (Strategy.validate); a C caller of the
    // reference ABI gets at least a warning: a rank outside a tree neither contributes to nor receives its slice.
    if (world > 0 && t.root >= 0 && (int)t.nodes.size() != world)
      ADAPCC_LOG(1, "strategy: tree %d covers %d of %d ranks", (int)trees.size(), (int)t.nodes.size(), world);
    if (t.root >= 0) trees.push_back(std::move(t));
    if ((int)trees.size() > kMaxTrees) { set_error("strategy: more than %d trees", kMaxTrees); return false; }
  }
  if (trees.empty()) { set_error("strategy: no usable <root> tree"); return false; }
  return true;
}

bool Strategy::load_file(const std::string& path, int world) {
  std::ifstream f(path);
  if (!f) { set_error("strategy: cannot open %s", path.c_str()); return false; }
  std::stringstream ss;
  ss << f.rdbuf();
  return load(ss.str(), world);
}

// ------------------------------------------------------------------------------------
// relay control
// ------------------------------------------------------------------------------------
static bool is_active(const std::vector<bool>& a, int r) { return r >= 0 && r < (int)a.size() && a[r]; }

static bool subtree_active(const StrategyTree& t, int x, const std::vector<bool>& a) {
  if (is_active(a, x)) return true;
  auto it = t.children.find(x);
  if (it == t.children.end()) return false;
  for (int c : it->second)
    if (subtree_active(t, c, a)) return true;
  return false;
}

RelayControl relay_control(const StrategyTree& tree, int rank, const std::vector<bool>& active) {
  RelayControl rc;
  auto it = tree.children.find(rank);
  if (it != tree.children.end())
    for (int c : it->second)
      if (subtree_active(tree, c, active)) rc.active_recvs.push_back(c);
  rc.has_recv = !rc.active_recvs.empty();
  rc.has_local = is_active(active, rank);
  // reference: no kernel when nothing arrives, or when a single flow merely passes through
  rc.has_kernel = rc.has_recv && !(rc.active_recvs.size() == 1 && !rc.has_local);
  const bool is_root = rank == tree.root;
  rc.has_send = (rc.has_local || rc.has_recv) && !is_root &&
                std::find(tree.nodes.begin(), tree.nodes.end(), rank) != tree.nodes.end();
  return rc;
}

HostTreeRole tree_role(const StrategyTree& tree_in, int rank, const std::vector<bool>& active,
                       int prim, int relay_mode) {
  HostTreeRole role;
  StrategyTree pruned;
  const StrategyTree* T = &tree_in;
  if (relay_mode == RELAY_BYPASS) {
    pruned = contract(tree_in, [&](int x) {
      // the broadcast root always stays: it owns the data
      return is_active(active, x) || (prim == BOARDCAST && x == tree_in.root);
    });
    T = &pruned;
  }
  if (std::find(T->nodes.begin(), T->nodes.end(), rank) == T->nodes.end()) return role;
  const bool is_root = rank == T->root;
  const bool local = is_active(active, rank);
  std::vector<int> recvs;
  auto it = T->children.find(rank);
  if (it != T->children.end())
    for (int c : it->second)
      if (subtree_active(*T, c, active)) recvs.push_back(c);
  auto pit = T->parent.find(rank);
  role.parent = is_root ? -1 : (pit == T->parent.end() ? -1 : pit->second);
  const int parent_is_root = (role.parent >= 0 && role.parent == T->root) ? TR_PARENT_IS_ROOT : 0;

  if (prim == ALLREDUCE || prim == REDUCE) {
    const bool in_reduce = local || !recvs.empty();
    if (!in_reduce) { role.parent = -1; return role; }
    role.flags |= TR_IN_REDUCE;
    if (local) role.flags |= TR_HAS_LOCAL;
    role.children = recvs;
    if (prim == ALLREDUCE) {
      if (!is_root) role.flags |= TR_IN_BCAST | parent_is_root;
      if (local) role.flags |= TR_WANT_RESULT;
      if (!recvs.empty()) role.flags |= TR_PUBLISH;
    } else if (is_root) {
      role.flags |= TR_WANT_RESULT;
    }
  } else if (prim == BOARDCAST) {
    const bool wants_below = !recvs.empty();
    if (is_root) {
      role.flags |= TR_PUBLISH;            // owns the data (never overwritten)
    } else {
      if (!local && !wants_below) { role.parent = -1; return role; }
      role.flags |= TR_IN_BCAST | parent_is_root;
      if (local) role.flags |= TR_WANT_RESULT;
      if (wants_below) role.flags |= TR_PUBLISH;
    }
  }
  return role;
}

}  // namespace adapcc
