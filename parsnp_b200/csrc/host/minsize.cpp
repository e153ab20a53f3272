#include "minsize.h"
#include <stdexcept>
#include <cmath>
#include <cctype>
#include <cstdlib>
#include <vector>

namespace pb200 {

namespace {
// fixed-capacity stack with the reference's (unchecked) semantics: pop on empty returns a default value
template <class T> struct Stk {
    std::vector<T> v;
    void push(T x) { v.push_back(x); }
    bool empty() const { return v.empty(); }
    T peek() const { return v.empty() ? T() : v.back(); }
    T pop() { if (v.empty()) return T(); T x = v.back(); v.pop_back(); return x; }
};
inline int prec_stop_muldiv(char top) { return top == '+' || top == '-'; }
}  // namespace

// Operator-precedence conversion with the reference's quirks (src/Converter.cpp:11-153):
//  * the expression is wrapped in one extra pair of parentheses and conversion stops when that closes;
//  * 'L' (from "Log") is pushed as a unary operator, the letters 'o','g' go to the output as operands;
//  * operands are emitted character by character, followed by a blank unless the next char continues a number.
std::string minsize_postfix(const std::string& infix_in) {
    std::string infix = infix_in + ")";
    std::string out;
    Stk<char> ops;
    ops.push('(');
    const int len = (int)infix.size();
    for (int i = 0; !ops.empty() && i < len; ++i) {
        const char ch = infix[i];
        if (ch == '(') { ops.push('('); continue; }
        if (ch == ')') {
            for (;;) { char b = ops.pop(); if (b == '(') break; out += b; if (ops.empty()) break; }
            continue;
        }
        if (ch == '+' || ch == '-') {
            if (ch == '-' && i == 0) { out += "-"; continue; }
            while (!ops.empty() && ops.peek() != '(') out += ops.pop();
            ops.push(ch);
            continue;
        }
        if (ch == '*' || ch == '/') {
            while (!ops.empty() && ops.peek() != '(' && !prec_stop_muldiv(ops.peek())) out += ops.pop();
            ops.push(ch);
            continue;
        }
        if (ch == '^') {
            while (!ops.empty() && ops.peek() == '^') out += ops.pop();
            ops.push('^');
            continue;
        }
        if (ch == 'L') {
            while (!ops.empty() && ops.peek() != '(') { ops.pop(); ops.pop(); }
            ops.push('L');
            continue;
        }
        if (ch == '\t' || ch == ' ') continue;
        if (isdigit((unsigned char)ch) || isalpha((unsigned char)ch)) {
            out += ch;
            const char nx = (i + 1 < len) ? infix[i + 1] : '\0';
            if (nx != '.' && !isdigit((unsigned char)nx)) out += ' ';
        }
        if (ch == '.') out += '.';
    }
    return out;
}

float minsize_eval(const std::string& in, float seqlen) {
    Stk<float> st;
    const int size = (int)in.size();
    int i = 0;
    while (i <= size) {
        const char a = (i < size) ? in[i] : '\0';
        if (isdigit((unsigned char)a) || a == '.') {
            const int start = i;
            while (i < size && isdigit((unsigned char)in[i])) {
                ++i;
                if (i < size && in[i] == '.') ++i;
            }
            float b = (float)atof(in.substr(start, i - start).c_str());
            st.push(b);
        } else {
            float x, y, z;
            switch (a) {
                case 'S': case 's': st.push(seqlen); break;
                case '+': x = st.pop(); y = st.pop(); z = y + x; st.push(z); break;
                case '-': x = st.pop(); y = st.pop(); z = y - x; st.push(z); break;
                case '*': x = st.pop(); y = st.pop(); z = y * x; st.push(z); break;
                case '/': x = st.pop(); y = st.pop(); if (x == 0) throw std::invalid_argument("parsnp_b200: division by zero in the minimum-length expression"); z = y / x; st.push(z); break;
                case '^': x = st.pop(); y = st.pop(); z = std::pow(y, x); st.push(z); break;
                case 'L': x = st.pop(); z = (float)(std::log(x) / std::log(2.0)); st.push(z); break;
                default: break;
            }
        }
        ++i;   // the reference advances once more after every token (skips the blank after a number)
    }
    float out = st.pop();
    return std::ceil(out);
}

int MinSizeExpr::operator()(int64_t slength) const {
    float limit = minsize_eval(postfix_, (float)slength);
    return (int)std::ceil(limit);
}

}  // namespace pb200
