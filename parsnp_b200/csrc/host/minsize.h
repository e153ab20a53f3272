// Minimum MUM length expression evaluator.
// Restates the observable behaviour of the reference's Converter()/Calculator() pair
// (src/Converter.cpp:11-153 infix->postfix, src/Converter.cpp:155-286 float RPN evaluation) as used by
// Aligner::setMums1 (src/parsnp.cpp:1502-1514):  minsize = int(ceil(Calculator(postfix(expr), S = slength))).
// All arithmetic is float32 exactly as in the reference (operands are `float`, Log is logf(x)/log(2.0) rounded to
// float) - this gates which MUMs exist, so it must be bit-exact.
#pragma once
#include <string>
#include <cstdint>

namespace pb200 {

// infix -> the reference's postfix string (e.g. "1.1*(Log(S))" -> "1.1 o g S L*")
std::string minsize_postfix(const std::string& infix);
// evaluate the postfix with S = seqlen (float), returns ceil()'d float like Calculator()
float minsize_eval(const std::string& postfix, float seqlen);

class MinSizeExpr {
public:
    explicit MinSizeExpr(const std::string& infix) : postfix_(minsize_postfix(infix)) {}
    int operator()(int64_t slength) const;   // int(ceil(limit))
    const std::string& postfix() const { return postfix_; }
private:
    std::string postfix_;
};

}  // namespace pb200
