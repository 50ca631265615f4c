// A small evaluator for the function expressions of WarpII input files.
//
// The reference hands "Function expression" strings to deal.II's Functions::ParsedFunction (muparser):
// src/five_moment/species_func.cc:32-51, e.g. `1 + 0.6 * sin(2*pi*x); 1.0; 0.0; 0.0; 1.0` or
// `if(x < 0.5, 1.0, 0.10)` (test/input_test.cc:32-34, 90-92).  This is an independent recursive-descent
// implementation of the subset those inputs use: numbers, named variables and constants, + - * / ^ (right
// associative, binds tighter than unary minus, as in muparser), comparisons, && ||, the ?: operator, and the usual
// functions including deal.II's if(c,a,b) and pow(a,b).  Components are separated by ';'.
// Like ParsedFunction::parse_parameters it defines the constants pi and Pi AFTER the user's list, so a user-supplied
// `pi=3.1415926535` is overridden by the exact value (the reference's GlobalIntegralsTest only passes that way).
#pragma once
#include <cmath>
#include <map>
#include <random>
#include <ctime>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace warpii_b200 {

class Expression {
   public:
    // variables: names bound positionally at evaluation time (e.g. {"x","y","t"})
    Expression(const std::string& text, const std::vector<std::string>& variables, const std::map<std::string, double>& constants)
        : src_(text), vars_(variables), consts_(constants) {
        pos_ = 0;
        root_ = parse_ternary();
        skip_ws();
        if (pos_ != src_.size()) fail("unexpected '" + src_.substr(pos_, 1) + "'");
    }
    double eval(const double* values) const { return eval_node(*root_, values); }
    // does the expression reference variable number `index`?
    bool uses_variable(int index) const { return uses(*root_, index); }

    // "a; b; c" -> {"a","b","c"}  (ParsedFunction component separator)
    static std::vector<std::string> split_components(const std::string& text) {
        std::vector<std::string> out;
        std::string cur;
        for (char ch : text) {
            if (ch == ';') { out.push_back(cur); cur.clear(); }
            else cur += ch;
        }
        out.push_back(cur);
        return out;
    }
    // "pi=3.14, k=2" -> map; then pi / Pi are set to the exact value (deal.II parsed_function.cc)
    static std::map<std::string, double> parse_constants(const std::string& text) {
        std::map<std::string, double> out;
        size_t i = 0;
        while (i < text.size()) {
            size_t j = text.find(',', i);
            if (j == std::string::npos) j = text.size();
            const std::string item = text.substr(i, j - i);
            const size_t eq = item.find('=');
            if (eq != std::string::npos) {
                const std::string name = trim(item.substr(0, eq));
                if (!name.empty()) out[name] = std::stod(item.substr(eq + 1));
            } else if (!trim(item).empty()) {
                throw std::invalid_argument("Function constants: expected name=value, got '" + item + "'");
            }
            i = j + 1;
        }
        out["pi"] = 3.14159265358979323846264338327950288;
        out["Pi"] = out["pi"];
        return out;
    }
    static std::string trim(const std::string& s) {
        size_t a = 0, b = s.size();
        while (a < b && std::isspace((unsigned char)s[a])) a++;
        while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
        return s.substr(a, b - a);
    }

   private:
    enum Kind { NUM, VAR, NEG, NOT, ADD, SUB, MUL, DIV, POW, LT, GT, LE, GE, EQ, NE, AND, OR, SEL, CALL1, CALL2 };
    struct Node {
        Kind kind;
        double value = 0;
        int index = 0;
        double (*f1)(double) = nullptr;
        double (*f2)(double, double) = nullptr;
        std::unique_ptr<Node> a, b, c;
    };
    using P = std::unique_ptr<Node>;

    [[noreturn]] void fail(const std::string& what) const {
        throw std::invalid_argument("expression '" + src_ + "': " + what + " at position " + std::to_string(pos_));
    }
    void skip_ws() { while (pos_ < src_.size() && std::isspace((unsigned char)src_[pos_])) pos_++; }
    bool eat(const char* tok) {
        skip_ws();
        const size_t n = std::char_traits<char>::length(tok);
        if (src_.compare(pos_, n, tok) == 0) { pos_ += n; return true; }
        return false;
    }
    static P make(Kind k, P a = nullptr, P b = nullptr, P c = nullptr) {
        P n(new Node());
        n->kind = k; n->a = std::move(a); n->b = std::move(b); n->c = std::move(c);
        return n;
    }

    P parse_ternary() {
        P cond = parse_or();
        if (eat("?")) {
            P x = parse_ternary();
            if (!eat(":")) fail("expected ':'");
            P y = parse_ternary();
            return make(SEL, std::move(cond), std::move(x), std::move(y));
        }
        return cond;
    }
    P parse_or() {
        P l = parse_and();
        while (eat("||") || eat("|")) l = make(OR, std::move(l), parse_and());
        return l;
    }
    P parse_and() {
        P l = parse_cmp();
        while (eat("&&") || eat("&")) l = make(AND, std::move(l), parse_cmp());
        return l;
    }
    P parse_cmp() {
        P l = parse_sum();
        for (;;) {
            if (eat("<=")) l = make(LE, std::move(l), parse_sum());
            else if (eat(">=")) l = make(GE, std::move(l), parse_sum());
            else if (eat("==")) l = make(EQ, std::move(l), parse_sum());
            else if (eat("!=")) l = make(NE, std::move(l), parse_sum());
            else if (eat("<")) l = make(LT, std::move(l), parse_sum());
            else if (eat(">")) l = make(GT, std::move(l), parse_sum());
            else return l;
        }
    }
    P parse_sum() {
        P l = parse_product();
        for (;;) {
            if (eat("+")) l = make(ADD, std::move(l), parse_product());
            else if (eat("-")) l = make(SUB, std::move(l), parse_product());
            else return l;
        }
    }
    P parse_product() {
        P l = parse_unary();
        for (;;) {
            if (eat("*")) l = make(MUL, std::move(l), parse_unary());
            else if (eat("/")) l = make(DIV, std::move(l), parse_unary());
            else return l;
        }
    }
    P parse_unary() {
        if (eat("-")) return make(NEG, parse_unary());   // -x^2 == -(x^2)
        if (eat("+")) return parse_unary();
        if (eat("!")) return make(NOT, parse_unary());
        return parse_power();
    }
    P parse_power() {
        P base = parse_atom();
        if (eat("^")) return make(POW, std::move(base), parse_unary());   // right associative
        return base;
    }
    P parse_atom() {
        skip_ws();
        if (pos_ >= src_.size()) fail("unexpected end");
        const char ch = src_[pos_];
        if (ch == '(') {
            pos_++;
            P e = parse_ternary();
            if (!eat(")")) fail("expected ')'");
            return e;
        }
        if (std::isdigit((unsigned char)ch) || ch == '.') {
            if (ch == '0' && pos_ + 1 < src_.size() && (src_[pos_ + 1] == 'x' || src_[pos_ + 1] == 'X')) fail("hexadecimal literals are not supported");
            size_t used = 0;
            double v = 0;
            try { v = std::stod(src_.substr(pos_), &used); } catch (...) { fail("bad number"); }
            pos_ += used;
            P n = make(NUM);
            n->value = v;
            return n;
        }
        if (std::isalpha((unsigned char)ch) || ch == '_') {
            size_t j = pos_;
            while (j < src_.size() && (std::isalnum((unsigned char)src_[j]) || src_[j] == '_')) j++;
            const std::string name = src_.substr(pos_, j - pos_);
            pos_ = j;
            skip_ws();
            if (pos_ < src_.size() && src_[pos_] == '(') {
                pos_++;
                std::vector<P> args;
                if (!eat(")")) {
                    do { args.push_back(parse_ternary()); } while (eat(","));
                    if (!eat(")")) fail("expected ')' after arguments of " + name);
                }
                return make_call(name, args);
            }
            for (size_t i = 0; i < vars_.size(); i++)
                if (vars_[i] == name) { P n = make(VAR); n->index = (int)i; return n; }
            auto it = consts_.find(name);
            if (it != consts_.end()) { P n = make(NUM); n->value = it->second; return n; }
            fail("unknown identifier '" + name + "'");
        }
        fail(std::string("unexpected '") + ch + "'");
    }
    P make_call(const std::string& name, std::vector<P>& args) {
        static const std::map<std::string, double (*)(double)> f1 = {
            {"sin", std::sin}, {"cos", std::cos}, {"tan", std::tan}, {"asin", std::asin}, {"acos", std::acos}, {"atan", std::atan},
            {"sinh", std::sinh}, {"cosh", std::cosh}, {"tanh", std::tanh}, {"exp", std::exp}, {"log", std::log}, {"ln", std::log},
            {"log2", std::log2}, {"log10", std::log10}, {"sqrt", std::sqrt}, {"abs", std::fabs}, {"ceil", std::ceil}, {"floor", std::floor},
            {"erfc", std::erfc}, {"erf", std::erf}, {"asinh", std::asinh}, {"acosh", std::acosh}, {"atanh", std::atanh}, {"rint", std::rint}, {"sign", [](double v) { return v < 0 ? -1.0 : (v > 0 ? 1.0 : 0.0); }},
            {"cot", [](double v) { return 1.0 / std::tan(v); }}, {"sec", [](double v) { return 1.0 / std::cos(v); }},
            {"csc", [](double v) { return 1.0 / std::sin(v); }}, {"int", [](double v) { return std::round(v); }},
            // deal.II's additions to muparser (function_parser / mu_rand_seed): a uniform [0,1) stream per seed value
            {"rand_seed", [](double seed) {
                 static std::map<double, std::mt19937> streams;
                 auto it = streams.find(seed);
                 if (it == streams.end()) it = streams.emplace(seed, std::mt19937((unsigned int)seed)).first;
                 return (double)it->second() / 4294967296.0;
             }}};
        if (name == "rand") {   // mu_rand: no argument, a stream seeded from the clock
            if (!args.empty()) fail("rand() takes no argument");
            P zero = make(NUM);
            zero->value = 0.0;
            P n = make(CALL1, std::move(zero));
            n->f1 = [](double) {
                static std::mt19937 stream((unsigned int)std::time(nullptr));
                return (double)stream() / 4294967296.0;
            };
            return n;
        }
        static const std::map<std::string, double (*)(double, double)> f2 = {
            {"pow", std::pow}, {"min", [](double a, double b) { return a < b ? a : b; }}, {"max", [](double a, double b) { return a > b ? a : b; }},
            {"atan2", std::atan2}, {"fmod", std::fmod}};
        if (name == "if") {
            if (args.size() != 3) fail("if() takes 3 arguments");
            return make(SEL, std::move(args[0]), std::move(args[1]), std::move(args[2]));
        }
        if ((name == "min" || name == "max" || name == "sum" || name == "avg") && args.size() != 2) {
            // muparser's variadic forms: fold the arguments pairwise (avg = sum / n)
            if (args.empty()) fail(name + "() needs at least 1 argument");
            const size_t n_args = args.size();
            P acc = std::move(args[0]);
            for (size_t i = 1; i < n_args; i++) {
                if (name == "sum" || name == "avg") acc = make(ADD, std::move(acc), std::move(args[i]));
                else {
                    P c = make(CALL2, std::move(acc), std::move(args[i]));
                    c->f2 = f2.find(name)->second;
                    acc = std::move(c);
                }
            }
            if (name == "avg") {
                P cnt = make(NUM);
                cnt->value = (double)n_args;
                acc = make(DIV, std::move(acc), std::move(cnt));
            }
            return acc;
        }
        if (name == "sum" || name == "avg") {   // exactly two arguments
            P acc = make(ADD, std::move(args[0]), std::move(args[1]));
            if (name == "avg") {
                P cnt = make(NUM);
                cnt->value = 2.0;
                acc = make(DIV, std::move(acc), std::move(cnt));
            }
            return acc;
        }
        auto i1 = f1.find(name);
        if (i1 != f1.end()) {
            if (args.size() != 1) fail(name + "() takes 1 argument");
            P n = make(CALL1, std::move(args[0]));
            n->f1 = i1->second;
            return n;
        }
        auto i2 = f2.find(name);
        if (i2 != f2.end()) {
            if (args.size() != 2) fail(name + "() takes 2 arguments");
            P n = make(CALL2, std::move(args[0]), std::move(args[1]));
            n->f2 = i2->second;
            return n;
        }
        fail("unknown function '" + name + "'");
    }

    static bool uses(const Node& n, int index) {
        if (n.kind == VAR && n.index == index) return true;
        return (n.a && uses(*n.a, index)) || (n.b && uses(*n.b, index)) || (n.c && uses(*n.c, index));
    }
    static double eval_node(const Node& n, const double* v) {
        switch (n.kind) {
            case NUM: return n.value;
            case VAR: return v[n.index];
            case NEG: return -eval_node(*n.a, v);
            case NOT: return eval_node(*n.a, v) == 0.0 ? 1.0 : 0.0;
            case ADD: return eval_node(*n.a, v) + eval_node(*n.b, v);
            case SUB: return eval_node(*n.a, v) - eval_node(*n.b, v);
            case MUL: return eval_node(*n.a, v) * eval_node(*n.b, v);
            case DIV: return eval_node(*n.a, v) / eval_node(*n.b, v);
            case POW: return std::pow(eval_node(*n.a, v), eval_node(*n.b, v));
            case LT: return eval_node(*n.a, v) < eval_node(*n.b, v) ? 1.0 : 0.0;
            case GT: return eval_node(*n.a, v) > eval_node(*n.b, v) ? 1.0 : 0.0;
            case LE: return eval_node(*n.a, v) <= eval_node(*n.b, v) ? 1.0 : 0.0;
            case GE: return eval_node(*n.a, v) >= eval_node(*n.b, v) ? 1.0 : 0.0;
            case EQ: return eval_node(*n.a, v) == eval_node(*n.b, v) ? 1.0 : 0.0;
            case NE: return eval_node(*n.a, v) != eval_node(*n.b, v) ? 1.0 : 0.0;
            case AND: return (eval_node(*n.a, v) != 0.0 && eval_node(*n.b, v) != 0.0) ? 1.0 : 0.0;
            case OR: return (eval_node(*n.a, v) != 0.0 || eval_node(*n.b, v) != 0.0) ? 1.0 : 0.0;
            case SEL: return eval_node(*n.a, v) != 0.0 ? eval_node(*n.b, v) : eval_node(*n.c, v);
            case CALL1: return n.f1(eval_node(*n.a, v));
            case CALL2: return n.f2(eval_node(*n.a, v), eval_node(*n.b, v));
        }
        return 0.0;
    }

    std::string src_;
    std::vector<std::string> vars_;
    std::map<std::string, double> consts_;
    size_t pos_ = 0;
    P root_;
};

}  // namespace warpii_b200
