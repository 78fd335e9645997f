// Minimal stand-in for <boost/program_options.hpp> covering exactly the calls in the reference's
// parse_args() (src/main.cpp:40-179). Oracle build only (test infrastructure, not product code).
#pragma once
#include <any>
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost { namespace program_options {

class error : public std::runtime_error { public: using std::runtime_error::runtime_error; };

class value_semantic {
public:
    virtual ~value_semantic() {}
    virtual bool is_switch() const = 0;
    virtual bool is_required() const = 0;
    virtual std::any parse(const std::string& name, const std::string& text) const = 0;
    virtual bool has_default(std::any&) const { return false; }
    virtual void notify(const std::any&) const = 0;
};

template <class T>
class typed_value : public value_semantic {
public:
    explicit typed_value(T* p, bool sw = false) : m_ptr(p), m_switch(sw) {}
    typed_value* required() { m_required = true; return this; }
    bool is_switch() const override { return m_switch; }
    bool is_required() const override { return m_required; }
    std::any parse(const std::string& name, const std::string& text) const override {
        if constexpr (std::is_same<T, std::string>::value) {
            return std::any(text);
        } else if constexpr (std::is_same<T, bool>::value) {
            return std::any(true);
        } else {
            std::istringstream is(text);
            T v{};
            if (std::is_unsigned<T>::value && !text.empty() && text[0] == '-')
                throw error("the argument ('" + text + "') for option '--" + name + "' is invalid");
            is >> v;
            if (is.fail() || !is.eof())
                throw error("the argument ('" + text + "') for option '--" + name + "' is invalid");
            return std::any(v);
        }
    }
    bool has_default(std::any& out) const override {
        if (m_switch) { out = std::any(T{}); return true; }
        return false;
    }
    void notify(const std::any& v) const override { if (m_ptr) *m_ptr = std::any_cast<T>(v); }
private:
    T* m_ptr;
    bool m_switch, m_required = false;
};

template <class T> typed_value<T>* value(T* p = nullptr) { return new typed_value<T>(p); }
inline typed_value<bool>* bool_switch(bool* p = nullptr) { return new typed_value<bool>(p, true); }

struct option_description {
    std::string long_name; char short_name = 0; std::string text;
    std::shared_ptr<const value_semantic> sem;   // null => flag without value ("help")
};

class options_description;
class options_description_easy_init {
public:
    explicit options_description_easy_init(options_description* o) : m_owner(o) {}
    options_description_easy_init& operator()(const char* name, const char* text);
    options_description_easy_init& operator()(const char* name, const value_semantic* s, const char* text);
private:
    options_description* m_owner;
};

class options_description {
public:
    explicit options_description(const std::string& caption) : m_caption(caption) {}
    options_description_easy_init add_options() { return options_description_easy_init(this); }
    void add(const char* name, const value_semantic* s, const char* text) {
        option_description d;
        std::string n(name);
        size_t c = n.find(',');
        if (c != std::string::npos) { d.long_name = n.substr(0, c); d.short_name = n[c + 1]; }
        else d.long_name = n;
        d.text = text; d.sem.reset(s);
        m_opts.push_back(d);
    }
    const std::vector<option_description>& options() const { return m_opts; }
    const std::string& caption() const { return m_caption; }
private:
    std::string m_caption;
    std::vector<option_description> m_opts;
};
inline options_description_easy_init& options_description_easy_init::operator()(const char* n, const char* t)
{ m_owner->add(n, nullptr, t); return *this; }
inline options_description_easy_init& options_description_easy_init::operator()(const char* n, const value_semantic* s, const char* t)
{ m_owner->add(n, s, t); return *this; }

inline std::ostream& operator<<(std::ostream& os, const options_description& d) {
    os << d.caption() << ":\n";
    for (auto& o : d.options()) {
        std::string head = "  ";
        if (o.short_name) { head += "-"; head += o.short_name; head += " [ --" + o.long_name + " ]"; }
        else head += "--" + o.long_name;
        if (o.sem && !o.sem->is_switch()) head += " arg";
        os << head << "\n        ";
        for (char c : o.text) { os << c; if (c == '\n') os << "        "; }
        os << "\n";
    }
    return os;
}

struct parsed_options {
    const options_description* desc;
    std::vector<std::pair<const option_description*, std::string>> items;
};

inline parsed_options parse_command_line(int argc, char** argv, const options_description& desc) {
    parsed_options out; out.desc = &desc;
    auto find_long = [&](const std::string& n) -> const option_description* {
        const option_description* hit = nullptr; int nhit = 0;
        for (auto& o : desc.options()) {
            if (o.long_name == n) return &o;
            if (o.long_name.compare(0, n.size(), n) == 0) { hit = &o; ++nhit; }
        }
        if (nhit == 1) return hit;
        if (nhit > 1) throw error("option '--" + n + "' is ambiguous");
        throw error("unrecognised option '--" + n + "'");
    };
    auto find_short = [&](char c) -> const option_description* {
        for (auto& o : desc.options()) if (o.short_name == c) return &o;
        throw error(std::string("unrecognised option '-") + c + "'");
    };
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        const option_description* o = nullptr;
        std::string val; bool has_val = false;
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            std::string n = a.substr(2);
            size_t eq = n.find('=');
            if (eq != std::string::npos) { val = n.substr(eq + 1); n = n.substr(0, eq); has_val = true; }
            o = find_long(n);
        } else if (a.size() >= 2 && a[0] == '-' && a[1] != '-') {
            o = find_short(a[1]);
            if (a.size() > 2) { val = a.substr(2); has_val = true; }
        } else {
            throw error("too many positional options have been specified on the command line");
        }
        bool wants = o->sem && !o->sem->is_switch();
        if (wants && !has_val) {
            if (i + 1 >= argc) throw error("the required argument for option '--" + o->long_name + "' is missing");
            val = argv[++i];
        }
        out.items.emplace_back(o, val);
    }
    return out;
}

class variable_value {
public:
    variable_value() {}
    explicit variable_value(std::any v) : m_v(std::move(v)) {}
    template <class T> const T& as() const { return *std::any_cast<T>(&m_v); }
    const std::any& raw() const { return m_v; }
private:
    std::any m_v;
};

class variables_map : public std::map<std::string, variable_value> {
public:
    size_t count(const std::string& k) const { return std::map<std::string, variable_value>::count(k); }
    const variable_value& operator[](const std::string& k) const {
        static variable_value empty;
        auto it = find(k);
        return it == end() ? empty : it->second;
    }
    const options_description* m_desc = nullptr;
};

inline void store(const parsed_options& p, variables_map& vm) {
    vm.m_desc = p.desc;
    for (auto& it : p.items) {
        const option_description* o = it.first;
        if (vm.count(o->long_name)) throw error("option '--" + o->long_name + "' cannot be specified more than once");
        std::any v = o->sem ? o->sem->parse(o->long_name, it.second) : std::any(std::string());
        vm.insert({o->long_name, variable_value(v)});
    }
    for (auto& o : p.desc->options()) {
        std::any d;
        if (o.sem && !vm.count(o.long_name) && o.sem->has_default(d)) vm.insert({o.long_name, variable_value(d)});
    }
}

inline void notify(variables_map& vm) {
    if (!vm.m_desc) return;
    for (auto& o : vm.m_desc->options()) {
        if (!o.sem) continue;
        if (!vm.count(o.long_name)) {
            if (o.sem->is_required()) throw error("the option '--" + o.long_name + "' is required but missing");
            continue;
        }
        o.sem->notify(vm[o.long_name].raw());
    }
}

}}  // namespace boost::program_options
