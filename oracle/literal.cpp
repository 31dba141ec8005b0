// oracle/literal.cpp — TEST INFRASTRUCTURE ONLY.  See literal.h for the reference citations.
#include "literal.h"

#include <algorithm>
#include <set>

namespace oracle {

using namespace gosyntax;

bool Seq::all_complete() const {
  if (lits.empty()) return false;  // reference literal/seq.go:206-215: empty Seq is not "all complete"
  for (auto& l : lits)
    if (!l.complete) return false;
  return true;
}

std::string Seq::longest_common_prefix() const {
  if (lits.empty()) return "";
  std::string p = lits[0].bytes;
  for (size_t i = 1; i < lits.size(); i++) {
    size_t k = 0;
    const std::string& b = lits[i].bytes;
    while (k < p.size() && k < b.size() && p[k] == b[k]) k++;
    p.resize(k);
    if (p.empty()) return "";
  }
  return p;
}

std::string Seq::longest_common_suffix() const {
  if (lits.empty()) return "";
  std::string s = lits[0].bytes;
  for (size_t i = 1; i < lits.size(); i++) {
    const std::string& b = lits[i].bytes;
    size_t k = 0;
    while (k < s.size() && k < b.size() && s[s.size() - 1 - k] == b[b.size() - 1 - k]) k++;
    s = s.substr(s.size() - k);
    if (s.empty()) return "";
  }
  return s;
}

void Seq::cross_forward(const Seq& other) {
  if (lits.empty() || other.lits.empty()) return;
  std::vector<Literal> out;
  for (auto& l : lits) {
    if (!l.complete) {
      out.push_back(l);
      continue;
    }
    for (auto& r : other.lits) out.push_back({l.bytes + r.bytes, r.complete});
  }
  lits.swap(out);
}

void Seq::keep_first_bytes(size_t n) {
  if (lits.empty() || n == 0) return;
  for (auto& l : lits)
    if (l.bytes.size() > n) {
      l.bytes.resize(n);
      l.complete = false;
    }
}

void Seq::dedup() {
  std::set<std::string> seen;
  std::vector<Literal> kept;
  for (auto& l : lits)
    if (seen.insert(l.bytes).second) kept.push_back(l);
  lits.swap(kept);
}

size_t Seq::min_len() const {
  size_t m = (size_t)-1;
  for (auto& l : lits) m = std::min(m, l.bytes.size());
  return m;
}

namespace {

std::string runesToBytes(const std::vector<int32_t>& rs) {
  std::string s;
  for (int32_t r : rs) {
    if (r < 0x80) {
      s += (char)r;
    } else if (r < 0x800) {
      s += (char)(0xC0 | (r >> 6));
      s += (char)(0x80 | (r & 0x3F));
    } else if (r < 0x10000) {
      s += (char)(0xE0 | (r >> 12));
      s += (char)(0x80 | ((r >> 6) & 0x3F));
      s += (char)(0x80 | (r & 0x3F));
    } else {
      s += (char)(0xF0 | (r >> 18));
      s += (char)(0x80 | ((r >> 12) & 0x3F));
      s += (char)(0x80 | ((r >> 6) & 0x3F));
      s += (char)(0x80 | (r & 0x3F));
    }
  }
  return s;
}

int32_t simpleFold(int32_t r) {
  if (r == 'K') return 'k';
  if (r == 'k') return 0x212A;
  if (r == 0x212A) return 'K';
  if (r == 'S') return 's';
  if (r == 's') return 0x17F;
  if (r == 0x17F) return 'S';
  if (r >= 'A' && r <= 'Z') return r + 32;
  if (r >= 'a' && r <= 'z') return r - 32;
  return r;
}

std::vector<int32_t> caseFolds(int32_t r) {
  std::vector<int32_t> out{r};
  for (int32_t f = simpleFold(r); f != r; f = simpleFold(f)) out.push_back(f);
  return out;
}

struct Extractor {
  ExtractorConfig cfg;

  Seq generateVariants(const std::vector<std::vector<int32_t>>& sets, size_t n) {
    std::vector<std::vector<int32_t>> variants{{}};
    for (size_t i = 0; i < n; i++) {
      std::vector<std::vector<int32_t>> next;
      for (auto& p : variants)
        for (int32_t r : sets[i]) {
          auto e = p;
          e.push_back(r);
          next.push_back(std::move(e));
        }
      variants.swap(next);
    }
    Seq s;
    for (auto& v : variants) {
      std::string b = runesToBytes(v);
      if ((int)b.size() > cfg.max_literal_len) b.resize(cfg.max_literal_len);
      s.lits.push_back({b, true});
    }
    return s;
  }

  Seq expandCaseFoldLiteral(const std::vector<int32_t>& runes) {
    if (runes.empty()) return Seq();
    int crossLimit = cfg.cross_product_limit > 0 ? cfg.cross_product_limit : 250;
    std::vector<std::vector<int32_t>> sets(runes.size());
    long total = 1;
    size_t filled = 0;
    for (size_t i = 0; i < runes.size(); i++) {
      sets[i] = caseFolds(runes[i]);
      filled = i + 1;
      total *= (long)sets[i].size();
      if (total > crossLimit) break;
    }
    if (total <= cfg.max_literals && filled == runes.size()) return generateVariants(sets, filled);
    // findMaxCaseFoldPrefix over ALL foldSets entries (unfilled ones are empty => product 0)
    size_t trim = sets.size();
    {
      long product = 1;
      for (size_t i = 0; i < sets.size(); i++) {
        product *= (long)sets[i].size();
        if (product > cfg.max_literals) {
          trim = i;
          break;
        }
      }
    }
    if (trim == 0) return Seq();
    // entries past `filled` are empty sets: generating variants over them yields nothing, as in Go
    Seq r = generateVariants(sets, trim);
    for (auto& l : r.lits) l.complete = false;
    r.dedup();
    if ((int)r.len() > cfg.max_literals) r.lits.resize(cfg.max_literals);
    return r;
  }

  Seq expandCharClass(const Regexp* re) {
    Seq out;
    if (re->op != OpCharClass) return out;
    long count = 0;
    for (size_t i = 0; i + 1 < re->rune.size(); i += 2) {
      count += re->rune[i + 1] - re->rune[i] + 1;
      if (count > cfg.max_class_size) return Seq();
    }
    for (size_t i = 0; i + 1 < re->rune.size(); i += 2)
      for (int32_t r = re->rune[i]; r <= re->rune[i + 1]; r++) {
        std::string b = runesToBytes({r});
        if ((int)b.size() > cfg.max_literal_len) b.resize(cfg.max_literal_len);
        out.lits.push_back({b, true});
        if ((int)out.len() >= cfg.max_literals) return out;
      }
    return out;
  }

  static void markAllInexact(Seq& s) {
    for (auto& l : s.lits) l.complete = false;
  }
  static bool hasAnyExact(const Seq& s) {
    for (auto& l : s.lits)
      if (l.complete) return true;
    return false;
  }

  Seq extractPrefixesAlternate(const Regexp* re, int depth) {
    int crossLimit = cfg.cross_product_limit > 0 ? cfg.cross_product_limit : 250;
    Seq result;
    bool overflowed = false;
    for (auto* sub : re->sub) {
      Seq s = extractPrefixes(sub, depth + 1);
      if (s.empty()) return Seq();
      for (auto& l : s.lits) {
        result.lits.push_back(l);
        if ((int)result.lits.size() > crossLimit) {
          overflowed = true;
          break;
        }
      }
      if (overflowed) break;
    }
    if (overflowed || (int)result.len() > cfg.max_literals) {
      result.keep_first_bytes(3);
      markAllInexact(result);
      result.dedup();
      if ((int)result.len() > cfg.max_literals) result.lits.resize(cfg.max_literals);
      if (overflowed) result.partial_coverage = true;
    }
    return result;
  }

  // returns false => "nil" (not expandable)
  bool expandAlternateContribution(const Regexp* alt, int depth, Seq& out) {
    if (alt->op != OpAlternate) return false;
    int crossLimit = cfg.cross_product_limit > 0 ? cfg.cross_product_limit : 250;
    std::vector<Literal> all;
    bool overflowed = false;
    for (auto* sub : alt->sub) {
      Seq s = extractPrefixes(sub, depth + 1);
      if (s.empty()) return false;
      if (overflowed) {
        for (auto& l : s.lits) {
          std::string b = l.bytes;
          if (b.size() > 3) b.resize(3);
          all.push_back({b, false});
        }
        if ((int)all.size() > crossLimit) {
          Seq t;
          t.lits = all;
          t.dedup();
          all = t.lits;
        }
        continue;
      }
      for (auto& l : s.lits) all.push_back(l);
      if ((int)all.size() > crossLimit) {
        overflowed = true;
        Seq t;
        t.lits = all;
        t.keep_first_bytes(3);
        markAllInexact(t);
        t.dedup();
        all = t.lits;
      }
    }
    out.lits = all;
    if (overflowed || (int)out.len() > cfg.max_literals) {
      out.keep_first_bytes(3);
      markAllInexact(out);
      out.dedup();
      if ((int)out.len() > cfg.max_literals) out.lits.resize(cfg.max_literals);
    }
    return true;
  }

  bool concatSubContribution(const Regexp* sub, int depth, Seq& out) {
    switch (sub->op) {
      case OpLiteral:
        if (sub->flags & FoldCase) {
          out = expandCaseFoldLiteral(sub->rune);
          return true;
        }
        out.lits = {{runesToBytes(sub->rune), true}};
        return true;
      case OpCharClass:
        out = expandCharClass(sub);
        return !out.empty();
      case OpAlternate:
        return expandAlternateContribution(sub, depth, out);
      case OpCapture:
        if (sub->sub.empty()) return false;
        return concatSubContribution(sub->sub[0], depth, out);
      case OpRepeat:
        if (sub->min >= 1 && !sub->sub.empty()) {
          if (!concatSubContribution(sub->sub[0], depth, out)) return false;
          for (auto& l : out.lits) l.complete = false;
          return true;
        }
        return false;
      case OpWordBoundary:
      case OpNoWordBoundary:
        out.lits = {{"", true}};
        return true;
      default:
        return false;
    }
  }

  Seq extractPrefixesConcat(const Regexp* re, int depth) {
    if (re->sub.empty()) return Seq();
    size_t start = 0;
    while (start < re->sub.size() &&
           (re->sub[start]->op == OpBeginLine || re->sub[start]->op == OpBeginText))
      start++;
    if (start >= re->sub.size()) return Seq();
    int crossLimit = cfg.cross_product_limit > 0 ? cfg.cross_product_limit : 250;
    Seq acc;
    acc.lits = {{"", true}};
    for (size_t i = start; i < re->sub.size(); i++) {
      if (!hasAnyExact(acc)) break;
      Seq contrib;
      if (!concatSubContribution(re->sub[i], depth, contrib)) {
        markAllInexact(acc);
        break;
      }
      acc.cross_forward(contrib);
      if ((int)acc.len() > crossLimit || (int)acc.len() > cfg.max_literals) {
        acc.keep_first_bytes(4);
        markAllInexact(acc);
        acc.dedup();
        if ((int)acc.len() > cfg.max_literals) acc.lits.resize(cfg.max_literals);
        break;
      }
      for (auto& l : acc.lits)
        if ((int)l.bytes.size() > cfg.max_literal_len) {
          l.bytes.resize(cfg.max_literal_len);
          l.complete = false;
        }
    }
    if (acc.len() == 1 && acc.lits[0].bytes.empty()) return Seq();
    return acc;
  }

  Seq extractPrefixes(const Regexp* re, int depth) {
    if (depth > 100) return Seq();
    switch (re->op) {
      case OpLiteral: {
        if (re->flags & FoldCase) return expandCaseFoldLiteral(re->rune);
        std::string b = runesToBytes(re->rune);
        if ((int)b.size() > cfg.max_literal_len) b.resize(cfg.max_literal_len);
        Seq s;
        s.lits = {{b, true}};
        return s;
      }
      case OpConcat: return extractPrefixesConcat(re, depth);
      case OpAlternate: return extractPrefixesAlternate(re, depth);
      case OpCharClass: return expandCharClass(re);
      case OpCapture:
        if (re->sub.empty()) return Seq();
        return extractPrefixes(re->sub[0], depth + 1);
      default:
        return Seq();
    }
  }

  Seq extractSuffixes(const Regexp* re, int depth) {
    if (depth > 100) return Seq();
    switch (re->op) {
      case OpLiteral: {
        if (re->flags & FoldCase) return expandCaseFoldLiteral(re->rune);
        std::string b = runesToBytes(re->rune);
        if ((int)b.size() > cfg.max_literal_len) b = b.substr(b.size() - cfg.max_literal_len);
        Seq s;
        s.lits = {{b, true}};
        return s;
      }
      case OpConcat: {
        if (re->sub.empty()) return Seq();
        int last = (int)re->sub.size() - 1;
        while (last >= 0) {
          Op op = re->sub[last]->op;
          if (op != OpEndLine && op != OpEndText && op != OpWordBoundary && op != OpNoWordBoundary)
            break;
          last--;
        }
        if (last < 0) return Seq();
        Seq suf = extractSuffixes(re->sub[last], depth + 1);
        if (suf.empty()) return Seq();
        for (int i = last - 1; i >= 0; i--) {
          const Regexp* sub = re->sub[i];
          if (sub->op == OpWordBoundary || sub->op == OpNoWordBoundary) continue;
          if (sub->op != OpLiteral) {
            for (auto& l : suf.lits) l.complete = false;
            return suf;
          }
          std::string prefix = runesToBytes(sub->rune);
          for (auto& l : suf.lits) {
            std::string nb = prefix + l.bytes;
            if ((int)nb.size() > cfg.max_literal_len) nb = nb.substr(nb.size() - cfg.max_literal_len);
            l.bytes = nb;
          }
          if ((int)suf.len() > cfg.max_literals) return suf;
        }
        return suf;
      }
      case OpAlternate: {
        Seq all;
        for (auto* sub : re->sub) {
          Seq s = extractSuffixes(sub, depth + 1);
          if (s.empty()) return Seq();
          for (auto& l : s.lits) {
            all.lits.push_back(l);
            if ((int)all.len() >= cfg.max_literals) return all;
          }
        }
        return all;
      }
      case OpCharClass: return expandCharClass(re);
      case OpCapture:
        if (re->sub.empty()) return Seq();
        return extractSuffixes(re->sub[0], depth + 1);
      default:
        return Seq();
    }
  }
};

}  // namespace

Seq ExtractPrefixes(const Regexp* re, const ExtractorConfig& cfg) {
  Extractor e{cfg};
  Seq seq = e.extractPrefixes(re, 0);
  // reference literal/extractor.go:135-149: cascade 4/3/2-byte trims to get back under 64
  if (seq.len() > 64) {
    Seq original = seq;
    const int attempts[3][2] = {{4, 64}, {3, 64}, {2, 64}};
    for (auto& a : attempts) {
      if ((int)seq.len() <= a[1]) break;
      seq.keep_first_bytes(a[0]);
      seq.dedup();
    }
    if (seq.len() > 64) seq = original;
  }
  return seq;
}

Seq ExtractSuffixes(const Regexp* re, const ExtractorConfig& cfg) {
  Extractor e{cfg};
  return e.extractSuffixes(re, 0);
}

}  // namespace oracle
