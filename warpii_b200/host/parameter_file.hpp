// Reader for WarpII input files.
//
// The reference parses its inputs with deal.II's ParameterHandler (warpii.cc:139-163, five_moment.cc:13-36): entries are
// declared with a default and a pattern, then `set name = value` lines inside `subsection NAME ... end` blocks are
// matched against the declarations.  This is an independent reader of that text format with the behaviour the
// reference relies on: '#' comments, '\' line continuation, nested subsections, declared entries with
// Integer/Double/Bool/Selection/MultipleSelection/Anything patterns, an error for undeclared entries unless the caller
// asks to skip them (the reference's first pass, parse_input_from_string(input, "", true)), and an error for a value
// that does not match its pattern.
#pragma once
#include <cmath>
#include <cstdlib>
#include <limits>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace warpii_b200 {

struct ParameterPattern {
    enum Kind { ANYTHING, INTEGER, DOUBLE, BOOL, SELECTION, MULTIPLE_SELECTION } kind = ANYTHING;
    double lo = -std::numeric_limits<double>::infinity(), hi = std::numeric_limits<double>::infinity();
    std::string options;   // "a|b|c"
    static ParameterPattern Anything() { return ParameterPattern(); }
    static ParameterPattern Integer(double lo = -std::numeric_limits<double>::infinity(), double hi = std::numeric_limits<double>::infinity()) {
        ParameterPattern p; p.kind = INTEGER; p.lo = lo; p.hi = hi; return p;
    }
    static ParameterPattern Double(double lo = -std::numeric_limits<double>::infinity(), double hi = std::numeric_limits<double>::infinity()) {
        ParameterPattern p; p.kind = DOUBLE; p.lo = lo; p.hi = hi; return p;
    }
    static ParameterPattern Bool() { ParameterPattern p; p.kind = BOOL; return p; }
    static ParameterPattern Selection(const std::string& options) { ParameterPattern p; p.kind = SELECTION; p.options = options; return p; }
    static ParameterPattern MultipleSelection(const std::string& options) { ParameterPattern p; p.kind = MULTIPLE_SELECTION; p.options = options; return p; }
};


class ParameterFile {
   public:
    using Pattern = ParameterPattern;

    void enter_subsection(const std::string& name) {
        path_.push_back(name);
        subsections_.insert(key_of(path_, ""));   // like ParameterHandler, entering a subsection declares it, even if it stays empty
    }
    void leave_subsection() {
        if (path_.empty()) throw std::logic_error("leave_subsection without enter_subsection");
        path_.pop_back();
    }

    void declare_entry(const std::string& name, const std::string& default_value, const Pattern& pattern = Pattern()) {
        const std::string key = key_of(path_, name);
        check_pattern(key, default_value, pattern);
        Entry& e = entries_[key];
        e.pattern = pattern;
        if (!e.declared) e.value = default_value;   // re-declaring keeps a value parsed earlier
        e.declared = true;
    }

    // skip_undefined: silently ignore entries and subsections that have not been declared
    void parse_input_from_string(const std::string& text, bool skip_undefined = false) {
        std::vector<std::string> saved = path_;
        std::istringstream in(text);
        std::string line, logical;
        int lineno = 0;
        size_t depth_at_start = path_.size();
        while (std::getline(in, line)) {
            lineno++;
            const size_t hash = line.find('#');
            if (hash != std::string::npos) line.erase(hash);
            line = trim(line);
            if (!line.empty() && line.back() == '\\') {   // continuation
                line.pop_back();
                logical += line + " ";
                continue;
            }
            logical += line;
            const std::string stmt = trim(logical);
            logical.clear();
            if (stmt.empty()) continue;
            try {
                statement(stmt, skip_undefined);
            } catch (const std::exception& e) {
                path_ = saved;
                throw std::invalid_argument("input line " + std::to_string(lineno) + ": " + e.what());
            }
        }
        if (path_.size() != depth_at_start) {
            path_ = saved;
            throw std::invalid_argument("input: unbalanced 'subsection'/'end'");
        }
    }

    std::string get(const std::string& name) const {
        auto it = entries_.find(key_of(path_, name));
        if (it == entries_.end() || !it->second.declared) throw std::invalid_argument("entry '" + key_of(path_, name) + "' has not been declared");
        return it->second.value;
    }
    long get_integer(const std::string& name) const { return std::strtol(get(name).c_str(), nullptr, 10); }
    double get_double(const std::string& name) const { return std::strtod(get(name).c_str(), nullptr); }
    bool get_bool(const std::string& name) const {
        const std::string v = get(name);
        return v == "true" || v == "yes" || v == "on";
    }

    // "0.0, 1.5" -> {0.0, 1.5}  (Patterns::Tools::Convert<Point<dim>> / std::array, grid_descriptions.cc:37-44)
    static std::vector<double> to_doubles(const std::string& s) {
        std::vector<double> out;
        for (const std::string& item : split(s, ',')) {
            const std::string t = trim(item);
            if (t.empty()) continue;
            char* end = nullptr;
            const double v = std::strtod(t.c_str(), &end);
            if (end == t.c_str() || *end != '\0') throw std::invalid_argument("'" + t + "' is not a number");
            out.push_back(v);
        }
        return out;
    }
    static std::vector<std::string> split(const std::string& s, char sep) {
        std::vector<std::string> out;
        std::string cur;
        for (char ch : s) {
            if (ch == sep) { out.push_back(cur); cur.clear(); }
            else cur += ch;
        }
        out.push_back(cur);
        return out;
    }
    static std::string trim(const std::string& s) {
        size_t a = 0, b = s.size();
        while (a < b && std::isspace((unsigned char)s[a])) a++;
        while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
        return s.substr(a, b - a);
    }

   private:
    struct Entry {
        std::string value;
        Pattern pattern;
        bool declared = false;
    };

    static std::string key_of(const std::vector<std::string>& path, const std::string& name) {
        std::string k;
        for (const auto& p : path) k += p + "/";
        return k + name;
    }
    bool subsection_declared(const std::string& prefix) const { return subsections_.count(prefix) != 0; }

    void statement(const std::string& stmt, bool skip_undefined) {
        if (starts_with_word(stmt, "subsection")) {
            const std::string name = trim(stmt.substr(10));
            if (name.empty()) throw std::invalid_argument("'subsection' without a name");
            path_.push_back(name);
            if (!skip_undefined && !subsection_declared(key_of(path_, ""))) {
                const std::string what = key_of(path_, "");
                path_.pop_back();
                throw std::invalid_argument("no subsection '" + what + "' has been declared");
            }
            return;
        }
        if (stmt == "end" || starts_with_word(stmt, "end")) {
            if (path_.empty()) throw std::invalid_argument("'end' without 'subsection'");
            path_.pop_back();
            return;
        }
        if (starts_with_word(stmt, "set")) {
            const size_t eq = stmt.find('=');
            if (eq == std::string::npos) throw std::invalid_argument("'set' without '='");
            const std::string name = trim(stmt.substr(3, eq - 3));
            const std::string value = trim(stmt.substr(eq + 1));
            const std::string key = key_of(path_, name);
            auto it = entries_.find(key);
            if (it == entries_.end() || !it->second.declared) {
                if (skip_undefined) return;   // the caller declares more entries and parses the text again
                throw std::invalid_argument("no entry with name '" + key + "' has been declared");
            }
            check_pattern(key, value, it->second.pattern);
            it->second.value = value;
            return;
        }
        throw std::invalid_argument("could not parse '" + stmt + "'");
    }

    static bool starts_with_word(const std::string& s, const std::string& w) {
        return s.compare(0, w.size(), w) == 0 && (s.size() == w.size() || std::isspace((unsigned char)s[w.size()]));
    }

    static void check_pattern(const std::string& key, const std::string& value, const Pattern& p) {
        auto bad = [&](const std::string& why) {
            throw std::invalid_argument("the value '" + value + "' of entry '" + key + "' " + why);
        };
        switch (p.kind) {
            case Pattern::ANYTHING: return;
            case Pattern::INTEGER: {
                char* end = nullptr;
                const long v = std::strtol(value.c_str(), &end, 10);
                if (end == value.c_str() || *end != '\0') bad("is not an integer");
                if ((double)v < p.lo || (double)v > p.hi) bad("is outside the allowed range");
                return;
            }
            case Pattern::DOUBLE: {
                char* end = nullptr;
                const double v = std::strtod(value.c_str(), &end);
                if (end == value.c_str() || *end != '\0') bad("is not a floating-point number");
                if (v < p.lo || v > p.hi) bad("is outside the allowed range");
                return;
            }
            case Pattern::BOOL:
                if (value != "true" && value != "false" && value != "yes" && value != "no" && value != "on" && value != "off") bad("is not a boolean");
                return;
            case Pattern::SELECTION: {
                for (const std::string& o : split(p.options, '|'))
                    if (o == value) return;
                bad("is not one of " + p.options);
                return;
            }
            case Pattern::MULTIPLE_SELECTION: {
                const std::vector<std::string> opts = split(p.options, '|');
                for (const std::string& item : split(value, ',')) {
                    const std::string t = trim(item);
                    if (t.empty()) continue;
                    bool ok = false;
                    for (const std::string& o : opts) ok = ok || o == t;
                    if (!ok) bad("contains '" + t + "', which is not one of " + p.options);
                }
                return;
            }
        }
    }

    std::vector<std::string> path_;
    std::map<std::string, Entry> entries_;
    std::set<std::string> subsections_;
};

}  // namespace warpii_b200
