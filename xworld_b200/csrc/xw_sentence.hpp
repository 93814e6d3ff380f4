// xw_sentence.hpp -- the teacher's language channel for the navigation tasks (host side).
//
// Replaces (reference file:line):
//   CFG.generate / RHS.value            python/context_free_grammar.py:41-49,166-188  (leftmost depth-first expansion,
//                                        random.choice over the alternatives of an unbound nonterminal)
//   the grammars and bindings of        games/xworld3d/tasks/XWorld3DNavTarget*.py (_define_grammar, idle: _bind ... _generate)
//                                        games/xworld/tasks/XWorldNav{Target,ColorTarget,Near,Between}.py
//   XWorldTask.simple_navigation_reward games/xworld/tasks/xworld_task.py:184-223 ("finish" / "timeup" sentences)
//
// The reference draws the productions from Python's unseeded `random`; here draw i of a sentence is Philox
// (seed, global env id, episode, site XW_SITE_SENTENCE, i) like every other reference `random` call site
// (xw_common.cuh), so a sentence is a pure function of the env's identity and state.  Parity with the reference is
// therefore membership: every sentence is one the reference grammar generates for the same bindings
// (tests/test_sentences.py checks it against the reference's own CFG class).
//
// Grammar notation (the reference's): `LHS -> alt | alt`, symbols separated by blanks, terminals in single quotes;
// `-->` marks a nonterminal that must be bound before generation.
#pragma once
#include <stdint.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "xw_common.cuh"

enum { XW_SITE_SENTENCE = 11 };

namespace xw_sentence {

struct Grammar {
    std::map<std::string, std::vector<std::string>> rules;
    void add(const std::string& lhs, const std::string& rhs) {  // rhs: alternatives separated by '|'
        std::vector<std::string>& v = rules[lhs];
        v.clear();
        size_t a = 0;
        while (a <= rhs.size()) {
            size_t b = rhs.find('|', a);
            if (b == std::string::npos) b = rhs.size();
            std::string alt = rhs.substr(a, b - a);
            size_t l = alt.find_first_not_of(' '), r = alt.find_last_not_of(' ');
            if (l != std::string::npos) v.push_back(alt.substr(l, r - l + 1));
            a = b + 1;
        }
    }
    bool bind(const std::string& lhs, const std::string& alt) {  // CFG.bind: narrow a rule down to one alternative
        auto it = rules.find(lhs);
        if (it == rules.end()) return false;
        for (const std::string& a : it->second)
            if (a == alt) { it->second.assign(1, alt); return true; }
        return false;
    }
};

// One sentence: leftmost depth-first expansion; draw() returns a uniform 32-bit word per call.
template <class Draw>
inline bool generate(const Grammar& g, const std::string& symbol, Draw& draw, std::string* out, int depth = 0) {
    if (symbol.size() >= 2 && symbol.front() == '\'' && symbol.back() == '\'') {
        if (!out->empty()) out->push_back(' ');
        out->append(symbol, 1, symbol.size() - 2);
        return true;
    }
    auto it = g.rules.find(symbol);
    if (it == g.rules.end() || it->second.empty() || depth > 16) return false;
    const std::vector<std::string>& alts = it->second;
    // random.choice(items): one draw even when a single alternative is left (RHS.value, context_free_grammar.py:41-49)
    const std::string& rhs = alts[xw_randbelow(draw(), (uint32_t)alts.size())];
    size_t a = 0;
    while (a < rhs.size()) {
        size_t b = rhs.find(' ', a);
        if (b == std::string::npos) b = rhs.size();
        if (b > a && !generate(g, rhs.substr(a, b - a), draw, out, depth + 1)) return false;
        a = b + 1;
    }
    return true;
}

// The grammar of one navigation task (rules: XW_RULES_*; task: XW_T3_* / XW_T2_*), goal names / colours / directions
// left unbound -- the caller binds G, G1, G2, O, T, C, D, P.  `names`, `colors`, `directions`: the alternatives of
// those slots ('a' | 'b' | ...), as XWorldTask._get_all_*_as_rhs builds them.
inline Grammar task_grammar(int rules, int task, const std::string& names, const std::string& colors, const std::string& directions) {
    Grammar g;
    const std::string go4 = "'go' 'to' | 'navigate' 'to' | 'reach' | 'move' 'to'", go5 = go4 + " | 'collect'";
    g.add("Y", "'Could' 'you' 'please' | 'Can' 'you' | 'Will' 'you'");
    g.add("timeup", "'Time' 'up' '.'");
    if (rules == XW_RULES_NAV3D) {
        g.add("S", "start | timeup | correct | wrong");
        g.add("correct", "'Well' 'done' '!'");
        g.add("wrong", "'Wrong' '!'");
        g.add("D", "'destination' | 'target' | 'goal' | 'end'");
        g.add("A", task == XW_T3_BETWEEN ? go4 : go5);
        switch (task) {
            case XW_T3_TARGET:
                g.add("start", "I0 | I1 | I2 | I3 | I4 | I5 | I6");
                g.add("I0", "G"); g.add("I1", "A G 'please' '.'"); g.add("I2", "'Please' A G '.'"); g.add("I3", "A G '.'");
                g.add("I4", "G 'is' 'your' D '.'"); g.add("I5", "G 'is' 'the' D '.'"); g.add("I6", "Y A G '?'");
                g.add("G", names);
                break;
            case XW_T3_AVOID:
                g.add("start", "I0 | I1 | I2 | I4 | I5 | I6");
                g.add("I0", "V G '.'"); g.add("I1", "V G 'please' '.'"); g.add("I2", "'Please' V G '.'");
                g.add("I4", "E G 'is' 'your' D '.'"); g.add("I5", "E G 'is' 'the' D '.'"); g.add("I6", "Y VV G '?'");
                g.add("V", "'do' 'not' A | 'avoid'"); g.add("VV", "'not' A | 'avoid'");
                g.add("E", "'anything' 'except' | 'anything' 'but'");
                g.add("G", names);
                break;
            case XW_T3_BETWEEN:
                g.add("start", "I0 | I1 | I2 | I3 | I4");
                g.add("I0", "A L B '.'"); g.add("I1", "A L B 'please' '.'"); g.add("I2", "'Please' A L B '.'");
                g.add("I3", "L B 'is' 'your' D '.'"); g.add("I4", "Y A L B '?'");
                g.add("B", "'between' G1 'and' G2");
                g.add("L", "'the' 'location' | 'the' 'grid' | 'the' 'place'");
                g.add("G1", names); g.add("G2", names);
                break;
            case XW_T3_DIRECTION:
                g.add("start", "I0 | I1 | I2 | I3 | I4");
                g.add("I0", "A NP G '.'"); g.add("I1", "A NP G 'please' '.'"); g.add("I2", "'Please' A NP G '.'");
                g.add("I3", "NP G 'is' 'your' D '.'"); g.add("I4", "Y A NP G '?'");
                g.add("NP", "'the' 'object' P | 'the' 'object' 'that' 'is' P");
                g.add("P", "LEFT | RIGHT | BEHIND | FRONT");
                g.add("LEFT", "'left' 'of' | 'to' 'the' 'left' 'of'"); g.add("RIGHT", "'right' 'of' | 'to' 'the' 'right' 'of'");
                g.add("BEHIND", "'behind'"); g.add("FRONT", "'in' 'the' 'front' 'of' | 'front' 'of'");
                g.add("G", names);
                break;
            default:  // XW_T3_NEAR
                g.add("start", "I0 | I1 | I2 | I3 | I4");
                g.add("I0", "A NP G"); g.add("I1", "A NP G 'please' '.'"); g.add("I2", "'Please' A NP G '.'");
                g.add("I3", "NP G 'is' 'your' D '.'"); g.add("I4", "Y A NP G '?'");
                g.add("NP", "'the' 'object' N"); g.add("N", "'near' | 'by' | 'besides'");
                g.add("G", names);
                break;
        }
        return g;
    }
    // walls.json: XWorldNav{Target,Near,ColorTarget,Between}
    g.add("S", "start | finish | timeup");
    g.add("finish", "'Well' 'done' '!'");
    g.add("A", go4);
    const char* dest = task == XW_T2_NEAR ? "dest" : "D";
    g.add(dest, "'destination' | 'target' | 'goal'");
    g.add("I1", "A G 'please' '.'"); g.add("I2", "'Please' A G '.'"); g.add("I3", "A G '.'");
    g.add("I4", std::string("G 'is' 'your' ") + dest + " '.'"); g.add("I5", std::string("G 'is' 'the' ") + dest + " '.'");
    g.add("I6", "Y A G '?'");
    if (task == XW_T2_TARGET) {
        g.add("start", "I1 | I2 | I3 | I4 | I5 | I6");
        g.add("G", names);
    } else {
        g.add("start", "I1 | I2 | I3 | I4 | I5 | I6 | I7");
        g.add("I7", "G '.'");
        if (task == XW_T2_COLOR_TARGET) { g.add("G", "C O"); g.add("C", colors); g.add("O", names); }
        else if (task == XW_T2_NEAR) { g.add("G", "D R O"); g.add("D", directions); g.add("R", "'to' | 'of' | 'near' | 'by'"); g.add("O", names); }
        else { g.add("G", "'the' 'grid' 'between' O 'and' T"); g.add("O", names); g.add("T", names); }
    }
    return g;
}

inline std::string quoted(const char* s) { return std::string("'") + (s ? s : "") + "'"; }

}  // namespace xw_sentence
