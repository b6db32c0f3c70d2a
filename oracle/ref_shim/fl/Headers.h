// Stand-in for fuzzylite 6 (fl/Headers.h): an independent, from-scratch mini-implementation of the published
// semantics the reference's src/fuzz/*.cpp call: linguistic variables with Trapezoid / Triangle terms, Mamdani rule
// blocks parsed from text ("if A is x and (B is y or B is z) then C is w"), Minimum / Maximum / AlgebraicSum /
// AlgebraicProduct norms, General activation, Maximum aggregation, Centroid defuzzifier, macheps comparisons.
// TEST INFRASTRUCTURE (oracle/_ref build only). It is a generic interpreter, deliberately structured differently from
// the closed-form evaluation in oracle/hmp_oracle.cpp, so that the two act as a cross-check of each other.
#pragma once
#include <cmath>
#include <cstdlib>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#define FL_AT __FILE__, __LINE__, __FUNCTION__

namespace fl {
typedef double scalar;
const scalar nan = std::numeric_limits<scalar>::quiet_NaN();
const scalar inf = std::numeric_limits<scalar>::infinity();
const std::nullptr_t null = nullptr;

class Exception : public std::exception {
public:
	explicit Exception(const std::string& what, const char* = "", int = 0, const char* = "") : what_(what) {}
	const char* what() const noexcept override { return what_.c_str(); }
	std::string getWhat() const { return what_; }
private:
	std::string what_;
};

class fuzzylite {
public:
	static scalar macheps() { return 1e-6; }
	static void setLogging(bool) {}
	static void setDebugging(bool) {}
};

struct Op {
	static bool isNaN(scalar x) { return x != x; }
	static bool isFinite(scalar x) { return !(x != x || x == inf || x == -inf); }
	static bool isEq(scalar a, scalar b, scalar eps = fuzzylite::macheps()) {
		return a == b || std::abs(a - b) < eps || (a != a && b != b);
	}
	static bool isLt(scalar a, scalar b, scalar eps = fuzzylite::macheps()) { return !isEq(a, b, eps) && a < b; }
	static bool isLE(scalar a, scalar b, scalar eps = fuzzylite::macheps()) { return isEq(a, b, eps) || a < b; }
	static bool isGt(scalar a, scalar b, scalar eps = fuzzylite::macheps()) { return !isEq(a, b, eps) && a > b; }
	static bool isGE(scalar a, scalar b, scalar eps = fuzzylite::macheps()) { return isEq(a, b, eps) || a > b; }
	static scalar min(scalar a, scalar b) {
		if (isNaN(a)) return b;
		if (isNaN(b)) return a;
		return a < b ? a : b;
	}
	static scalar max(scalar a, scalar b) {
		if (isNaN(a)) return b;
		if (isNaN(b)) return a;
		return a > b ? a : b;
	}
	static scalar bound(scalar x, scalar lo, scalar hi) {
		if (x > hi) return hi;
		if (x < lo) return lo;
		return x;
	}
	static scalar toScalar(const std::string& s) {
		if (s == "nan" || s == "-nan" || s == "NaN") return nan;
		if (s == "inf" || s == "+inf") return inf;
		if (s == "-inf") return -inf;
		char* end = nullptr;
		scalar v = std::strtod(s.c_str(), &end);
		if (end == s.c_str() || *end != '\0') throw Exception("[conversion error] from <" + s + "> to scalar");
		return v;
	}
	static std::vector<std::string> split(const std::string& s) {
		std::vector<std::string> out;
		std::istringstream is(s);
		std::string tok;
		while (is >> tok) out.push_back(tok);
		return out;
	}
	static std::string str(scalar x) {
		std::ostringstream os;
		os.setf(std::ios::fixed);
		os.precision(3);
		os << x;
		return os.str();
	}
};

// ---- norms --------------------------------------------------------------------------------------
class Norm {
public:
	virtual ~Norm() {}
	virtual scalar compute(scalar a, scalar b) const = 0;
};
class TNorm : public Norm {};
class SNorm : public Norm {};
class Minimum : public TNorm {
public:
	scalar compute(scalar a, scalar b) const override { return Op::min(a, b); }
};
class AlgebraicProduct : public TNorm {
public:
	scalar compute(scalar a, scalar b) const override { return a * b; }
};
class Maximum : public SNorm {
public:
	scalar compute(scalar a, scalar b) const override { return Op::max(a, b); }
};
class AlgebraicSum : public SNorm {
public:
	scalar compute(scalar a, scalar b) const override { return a + b - (a * b); }
};

// ---- terms --------------------------------------------------------------------------------------
class Term {
public:
	explicit Term(const std::string& name = "", scalar height = 1.0) : name_(name), height_(height) {}
	virtual ~Term() {}
	virtual scalar membership(scalar x) const = 0;
	virtual void configure(const std::string& parameters) = 0;
	void setName(const std::string& n) { name_ = n; }
	std::string getName() const { return name_; }
	void setHeight(scalar h) { height_ = h; }
	scalar getHeight() const { return height_; }
protected:
	std::string name_;
	scalar height_;
};

class Trapezoid : public Term {
public:
	explicit Trapezoid(const std::string& name = "", scalar a = nan, scalar b = nan, scalar c = nan, scalar d = nan,
	                   scalar height = 1.0)
	    : Term(name, height), a_(a), b_(b), c_(c), d_(d) {}
	scalar membership(scalar x) const override {
		if (Op::isNaN(x)) return nan;
		if (Op::isLt(x, a_) || Op::isGt(x, d_)) return height_ * 0.0;
		if (Op::isLt(x, b_)) return height_ * Op::min(scalar(1.0), (x - a_) / (b_ - a_));
		if (Op::isLE(x, c_)) return height_ * 1.0;
		if (Op::isLt(x, d_)) return height_ * (d_ - x) / (d_ - c_);
		if (d_ == inf) return height_ * 1.0;
		return height_ * 0.0;
	}
	void configure(const std::string& parameters) override {
		if (parameters.empty()) return;
		std::vector<std::string> v = Op::split(parameters);
		if (v.size() < 4) throw Exception("[configuration error] term <Trapezoid> requires <4> parameters");
		a_ = Op::toScalar(v[0]);
		b_ = Op::toScalar(v[1]);
		c_ = Op::toScalar(v[2]);
		d_ = Op::toScalar(v[3]);
		if (v.size() > 4) height_ = Op::toScalar(v[4]);
	}
	void setVertexA(scalar v) { a_ = v; }
	void setVertexB(scalar v) { b_ = v; }
	void setVertexC(scalar v) { c_ = v; }
	void setVertexD(scalar v) { d_ = v; }
	scalar getVertexA() const { return a_; }
	scalar getVertexB() const { return b_; }
	scalar getVertexC() const { return c_; }
	scalar getVertexD() const { return d_; }
private:
	scalar a_, b_, c_, d_;
};

class Triangle : public Term {
public:
	explicit Triangle(const std::string& name = "", scalar a = nan, scalar b = nan, scalar c = nan, scalar height = 1.0)
	    : Term(name, height), a_(a), b_(b), c_(c) {}
	scalar membership(scalar x) const override {
		if (Op::isNaN(x)) return nan;
		if (Op::isLt(x, a_) || Op::isGt(x, c_)) return height_ * 0.0;
		if (Op::isEq(x, b_)) return height_ * 1.0;
		if (Op::isLt(x, b_)) {
			if (a_ == -inf) return height_ * 1.0;
			return height_ * (x - a_) / (b_ - a_);
		}
		if (c_ == inf) return height_ * 1.0;
		return height_ * (c_ - x) / (c_ - b_);
	}
	void configure(const std::string& parameters) override {
		if (parameters.empty()) return;
		std::vector<std::string> v = Op::split(parameters);
		if (v.size() < 3) throw Exception("[configuration error] term <Triangle> requires <3> parameters");
		a_ = Op::toScalar(v[0]);
		b_ = Op::toScalar(v[1]);
		c_ = Op::toScalar(v[2]);
		if (v.size() > 3) height_ = Op::toScalar(v[3]);
	}
private:
	scalar a_, b_, c_;
};

// A term of the consequent, scaled by the rule's activation degree through the implication operator.
class Activated : public Term {
public:
	Activated(const Term* term, scalar degree, const TNorm* implication)
	    : Term(term->getName()), term_(term), degree_(degree), implication_(implication) {}
	scalar membership(scalar x) const override {
		if (Op::isNaN(x)) return nan;
		return implication_->compute(term_->membership(x), degree_);
	}
	void configure(const std::string&) override {}
	const Term* getTerm() const { return term_; }
	scalar getDegree() const { return degree_; }
private:
	const Term* term_;
	scalar degree_;
	const TNorm* implication_;
};

// The fuzzy set an output variable accumulates during Engine::process().
class Aggregated : public Term {
public:
	Aggregated() : Term("fuzzyOutput"), aggregation_(nullptr) {}
	scalar membership(scalar x) const override {
		if (Op::isNaN(x)) return nan;
		if (terms_.empty()) return 0.0;
		if (!aggregation_) throw Exception("[aggregation error] aggregation operator needed to aggregate variable");
		scalar mu = 0.0;
		for (const Activated& t : terms_) mu = aggregation_->compute(mu, t.membership(x));
		return mu;
	}
	void configure(const std::string&) override {}
	void addTerm(const Activated& t) { terms_.push_back(t); }
	void clear() { terms_.clear(); }
	bool isEmpty() const { return terms_.empty(); }
	void setAggregation(SNorm* a) { aggregation_.reset(a); }
	const std::vector<Activated>& terms() const { return terms_; }
private:
	std::vector<Activated> terms_;
	std::unique_ptr<SNorm> aggregation_;
};

// ---- defuzzifier --------------------------------------------------------------------------------
class Defuzzifier {
public:
	virtual ~Defuzzifier() {}
	virtual scalar defuzzify(const Term* term, scalar minimum, scalar maximum) const = 0;
};
class Centroid : public Defuzzifier {
public:
	explicit Centroid(int resolution = 100) : resolution_(resolution) {}
	scalar defuzzify(const Term* term, scalar minimum, scalar maximum) const override {
		if (!Op::isFinite(minimum + maximum)) return nan;
		const scalar dx = (maximum - minimum) / resolution_;
		scalar area = 0, xcentroid = 0;
		for (int i = 0; i < resolution_; ++i) {
			scalar x = minimum + (i + 0.5) * dx;
			scalar y = term->membership(x);
			xcentroid += y * x;
			area += y;
		}
		return xcentroid / area;
	}
private:
	int resolution_;
};

// ---- variables ----------------------------------------------------------------------------------
class Variable {
public:
	Variable() : value_(nan), min_(-inf), max_(inf), enabled_(true), lock_(false) {}
	virtual ~Variable() {
		for (Term* t : terms_) delete t;
	}
	void setName(const std::string& n) { name_ = n; }
	std::string getName() const { return name_; }
	void setDescription(const std::string&) {}
	void setEnabled(bool e) { enabled_ = e; }
	bool isEnabled() const { return enabled_; }
	void setRange(scalar lo, scalar hi) { min_ = lo; max_ = hi; }
	scalar getMinimum() const { return min_; }
	scalar getMaximum() const { return max_; }
	void setLockValueInRange(bool l) { lock_ = l; }
	void setValue(scalar v) { value_ = lock_ ? Op::bound(v, min_, max_) : v; }
	scalar getValue() const { return value_; }
	void addTerm(Term* t) { terms_.push_back(t); }
	const std::vector<Term*>& terms() const { return terms_; }
	Term* getTerm(const std::string& name) const {
		for (Term* t : terms_) {
			if (t->getName() == name) return t;
		}
		throw Exception("[variable error] term <" + name + "> not found in variable <" + name_ + ">");
	}
	Term* highestMembership(scalar x, scalar* yhighest = nullptr) const {
		Term* result = nullptr;
		scalar ymax = 0.0;
		for (Term* t : terms_) {
			scalar y = t->membership(x);
			if (Op::isGt(y, ymax)) {
				ymax = y;
				result = t;
			}
		}
		if (yhighest) *yhighest = ymax;
		return result;
	}
	std::string fuzzify(scalar x) const {
		std::ostringstream os;
		for (size_t i = 0; i < terms_.size(); ++i) {
			if (i) os << " + ";
			os << Op::str(terms_[i]->membership(x)) << "/" << terms_[i]->getName();
		}
		return os.str();
	}
protected:
	std::string name_;
	std::vector<Term*> terms_;
	scalar value_, min_, max_;
	bool enabled_, lock_;
};

class InputVariable : public Variable {
public:
	std::string fuzzyInputValue() const { return fuzzify(getValue()); }
};

class OutputVariable : public Variable {
public:
	OutputVariable() : default_(nan), previous_(nan), lock_previous_(false) {}
	Aggregated* fuzzyOutput() { return &fuzzy_; }
	void setAggregation(SNorm* a) { fuzzy_.setAggregation(a); }
	void setDefuzzifier(Defuzzifier* d) { defuzzifier_.reset(d); }
	void setDefaultValue(scalar v) { default_ = v; }
	void setLockPreviousValue(bool l) { lock_previous_ = l; }
	void defuzzify() {
		if (!isEnabled()) return;
		if (Op::isFinite(getValue())) previous_ = getValue();
		scalar result = nan;
		if (!fuzzy_.isEmpty()) {
			if (!defuzzifier_) throw Exception("[defuzzifier error] defuzzifier needed to defuzzify output variable");
			result = defuzzifier_->defuzzify(&fuzzy_, getMinimum(), getMaximum());
		} else if (lock_previous_ && !Op::isNaN(previous_)) {
			result = previous_;
		} else {
			result = default_;
		}
		setValue(result);
	}
	std::string fuzzyOutputValue() const {
		std::ostringstream os;
		bool first = true;
		for (Term* t : terms_) {
			scalar degree = 0.0;
			for (const Activated& a : fuzzy_.terms()) {
				if (a.getTerm() == t) degree = Op::max(degree, a.getDegree());
			}
			if (!first) os << " + ";
			first = false;
			os << Op::str(degree) << "/" << t->getName();
		}
		return os.str();
	}
private:
	Aggregated fuzzy_;
	std::unique_ptr<Defuzzifier> defuzzifier_;
	scalar default_, previous_;
	bool lock_previous_;
};

// ---- rules --------------------------------------------------------------------------------------
class Engine;

// Antecedent expression tree: leaf = "<variable> is <term>", inner node = and / or.
struct AntecedentNode {
	enum Kind { PROPOSITION, AND, OR } kind = PROPOSITION;
	const Variable* variable = nullptr;
	const Term* term = nullptr;
	std::unique_ptr<AntecedentNode> left, right;
	scalar degree(const TNorm* conjunction, const SNorm* disjunction) const {
		if (kind == PROPOSITION) {
			if (!variable->isEnabled()) return 0.0;
			return term->membership(variable->getValue());
		}
		scalar l = left->degree(conjunction, disjunction);
		scalar r = right->degree(conjunction, disjunction);
		if (kind == AND) {
			if (!conjunction) throw Exception("[conjunction error] the following rule requires a conjunction operator");
			return conjunction->compute(l, r);
		}
		if (!disjunction) throw Exception("[disjunction error] the following rule requires a disjunction operator");
		return disjunction->compute(l, r);
	}
};

class Rule {
public:
	static Rule* parse(const std::string& text, const Engine* engine);
	const std::string& getText() const { return text_; }
	void deactivate() { degree_ = 0.0; triggered_ = false; }
	scalar activateWith(const TNorm* conjunction, const SNorm* disjunction) {
		degree_ = weight_ * antecedent_->degree(conjunction, disjunction);
		return degree_;
	}
	void trigger(const TNorm* implication) {
		if (Op::isGt(degree_, 0.0)) {
			if (!implication) throw Exception("[implication error] implication operator needed");
			for (auto& c : consequent_) {
				if (c.first->isEnabled()) c.first->fuzzyOutput()->addTerm(Activated(c.second, degree_, implication));
			}
			triggered_ = true;
		}
	}
	scalar getActivationDegree() const { return degree_; }
	bool isTriggered() const { return triggered_; }
private:
	std::string text_;
	scalar weight_ = 1.0, degree_ = 0.0;
	bool triggered_ = false;
	std::unique_ptr<AntecedentNode> antecedent_;
	std::vector<std::pair<OutputVariable*, const Term*>> consequent_;
};

class Activation {
public:
	virtual ~Activation() {}
};
class General : public Activation {};

class RuleBlock {
public:
	~RuleBlock() {
		for (Rule* r : rules_) delete r;
	}
	void setName(const std::string& n) { name_ = n; }
	void setDescription(const std::string&) {}
	void setEnabled(bool e) { enabled_ = e; }
	bool isEnabled() const { return enabled_; }
	void setConjunction(TNorm* n) { conjunction_.reset(n); }
	void setDisjunction(SNorm* n) { disjunction_.reset(n); }
	void setImplication(TNorm* n) { implication_.reset(n); }
	void setActivation(Activation* a) { activation_.reset(a); }
	void addRule(Rule* r) { rules_.push_back(r); }
	// General activation: every rule in order
	void activate() {
		for (Rule* r : rules_) {
			r->deactivate();
			r->activateWith(conjunction_.get(), disjunction_.get());
			r->trigger(implication_.get());
		}
	}
	std::string toString() const {
		std::ostringstream os;
		os << "RuleBlock: " << name_ << "\n";
		for (Rule* r : rules_) os << "  rule: " << r->getText() << "\n";
		return os.str();
	}
	const std::vector<Rule*>& rules() const { return rules_; }
private:
	std::string name_;
	bool enabled_ = true;
	std::unique_ptr<TNorm> conjunction_, implication_;
	std::unique_ptr<SNorm> disjunction_;
	std::unique_ptr<Activation> activation_;
	std::vector<Rule*> rules_;
};

class Engine {
public:
	~Engine() {
		for (auto* v : inputs_) delete v;
		for (auto* v : outputs_) delete v;
		for (auto* b : blocks_) delete b;
	}
	void setName(const std::string& n) { name_ = n; }
	void setDescription(const std::string&) {}
	void addInputVariable(InputVariable* v) { inputs_.push_back(v); }
	void addOutputVariable(OutputVariable* v) { outputs_.push_back(v); }
	void addRuleBlock(RuleBlock* b) { blocks_.push_back(b); }
	size_t numberOfInputVariables() const { return inputs_.size(); }
	size_t numberOfOutputVariables() const { return outputs_.size(); }
	size_t numberOfRuleBlocks() const { return blocks_.size(); }
	bool isReady(std::string* status = nullptr) const {
		if (status) status->clear();
		return !inputs_.empty() && !outputs_.empty() && !blocks_.empty();
	}
	void process() {
		for (OutputVariable* o : outputs_) o->fuzzyOutput()->clear();
		for (RuleBlock* b : blocks_) {
			if (b->isEnabled()) b->activate();
		}
		for (OutputVariable* o : outputs_) o->defuzzify();
	}
	InputVariable* getInputVariable(const std::string& name) const {
		for (auto* v : inputs_) {
			if (v->getName() == name) return v;
		}
		throw Exception("[engine error] input variable <" + name + "> not found");
	}
	OutputVariable* getOutputVariable(const std::string& name) const {
		for (auto* v : outputs_) {
			if (v->getName() == name) return v;
		}
		throw Exception("[engine error] output variable <" + name + "> not found");
	}
private:
	std::string name_;
	std::vector<InputVariable*> inputs_;
	std::vector<OutputVariable*> outputs_;
	std::vector<RuleBlock*> blocks_;
};

// "if <antecedent> then <variable> is <term> [and <variable> is <term>]* [with <weight>]"; `and` binds tighter than `or`.
inline Rule* Rule::parse(const std::string& text, const Engine* engine) {
	std::string spaced;
	for (char c : text) {
		if (c == '(' || c == ')') {
			spaced += ' ';
			spaced += c;
			spaced += ' ';
		} else {
			spaced += c;
		}
	}
	std::vector<std::string> tok = Op::split(spaced);
	size_t i = 0;
	if (tok.empty() || tok[i++] != "if") throw Exception("[syntax error] rule does not start with 'if': " + text);
	std::unique_ptr<Rule> rule(new Rule);
	rule->text_ = text;
	// recursive descent: or_expr := and_expr ('or' and_expr)* ; and_expr := primary ('and' primary)* ;
	// primary := '(' or_expr ')' | <variable> 'is' <term>
	struct Parser {
		const std::vector<std::string>& t;
		size_t& i;
		const Engine* e;
		const std::string& text;
		std::unique_ptr<AntecedentNode> primary() {
			if (i >= t.size()) throw Exception("[syntax error] unexpected end of antecedent: " + text);
			if (t[i] == "(") {
				++i;
				auto n = orExpr();
				if (i >= t.size() || t[i] != ")") throw Exception("[syntax error] missing ')': " + text);
				++i;
				return n;
			}
			if (i + 2 >= t.size() || t[i + 1] != "is") throw Exception("[syntax error] expected '<variable> is <term>': " + text);
			std::unique_ptr<AntecedentNode> n(new AntecedentNode);
			n->variable = e->getInputVariable(t[i]);
			n->term = n->variable->getTerm(t[i + 2]);
			i += 3;
			return n;
		}
		std::unique_ptr<AntecedentNode> andExpr() {
			auto l = primary();
			while (i < t.size() && t[i] == "and") {
				++i;
				std::unique_ptr<AntecedentNode> n(new AntecedentNode);
				n->kind = AntecedentNode::AND;
				n->left = std::move(l);
				n->right = primary();
				l = std::move(n);
			}
			return l;
		}
		std::unique_ptr<AntecedentNode> orExpr() {
			auto l = andExpr();
			while (i < t.size() && t[i] == "or") {
				++i;
				std::unique_ptr<AntecedentNode> n(new AntecedentNode);
				n->kind = AntecedentNode::OR;
				n->left = std::move(l);
				n->right = andExpr();
				l = std::move(n);
			}
			return l;
		}
	} parser{tok, i, engine, text};
	rule->antecedent_ = parser.orExpr();
	if (i >= tok.size() || tok[i++] != "then") throw Exception("[syntax error] missing 'then': " + text);
	while (true) {
		if (i + 3 > tok.size() || tok[i + 1] != "is") throw Exception("[syntax error] expected '<variable> is <term>': " + text);
		OutputVariable* v = engine->getOutputVariable(tok[i]);
		rule->consequent_.emplace_back(v, v->getTerm(tok[i + 2]));
		i += 3;
		if (i < tok.size() && tok[i] == "and") {
			++i;
			continue;
		}
		break;
	}
	if (i < tok.size() && tok[i] == "with") {
		if (i + 1 >= tok.size()) throw Exception("[syntax error] missing weight: " + text);
		rule->weight_ = Op::toScalar(tok[i + 1]);
		i += 2;
	}
	if (i != tok.size()) throw Exception("[syntax error] unexpected token <" + tok[i] + ">: " + text);
	return rule.release();
}

}  // namespace fl
