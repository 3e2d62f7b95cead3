// Minimal stand-in for boost::program_options, written from scratch for the oracle build only.
// It implements just the surface the reference's Settings class touches
// (utils/settings.h:24-46, utils/settings.cpp:40-331,458-505): option groups with
// add_options()(...) chains, typed values with default_value(), one positional option,
// GNU-style long/short command-line parsing with unique-prefix guessing (README.md:37-39),
// "key = value" config files, store()/notify(), and typed reads through variable_value::as<T>().
// TEST INFRASTRUCTURE - never linked into the product library.
#ifndef PAGAN2_B200_SHIM_PROGRAM_OPTIONS_HPP
#define PAGAN2_B200_SHIM_PROGRAM_OPTIONS_HPP
#include <map>
#include <string>
#include <vector>
#include <sstream>
#include <istream>
#include <ostream>
#include <cstdlib>
#include <memory>
#include <stdexcept>

namespace boost { namespace program_options {

class error : public std::runtime_error {
public:
    explicit error(const std::string &w) : std::runtime_error(w) {}
};

// ---------------------------------------------------------------- values
class variable_value {
    std::string text_;
    bool defaulted_;
public:
    variable_value() : defaulted_(false) {}
    variable_value(const std::string &t, bool d) : text_(t), defaulted_(d) {}
    const std::string &text() const { return text_; }
    bool defaulted() const { return defaulted_; }
    bool empty() const { return text_.empty(); }
    template <class T> T as() const {
        std::istringstream is(text_);
        T v = T();
        is >> v;
        return v;
    }
};
template <> inline std::string variable_value::as<std::string>() const { return text_; }
template <> inline int variable_value::as<int>() const { return (int)std::strtol(text_.c_str(), 0, 10); }
template <> inline float variable_value::as<float>() const { return std::strtof(text_.c_str(), 0); }
template <> inline double variable_value::as<double>() const { return std::strtod(text_.c_str(), 0); }

class value_semantic {
public:
    bool has_default;
    std::string default_text;
    value_semantic() : has_default(false) {}
    virtual ~value_semantic() {}
};

template <class T> class typed_value : public value_semantic {
public:
    typed_value *default_value(const T &v) {
        std::ostringstream os;
        os.precision(9);
        os << v;
        has_default = true;
        default_text = os.str();
        return this;
    }
    typed_value *default_value(const T &v, const std::string &textual) {
        (void)textual;
        return default_value(v);
    }
};

template <class T> typed_value<T> *value() { return new typed_value<T>(); }

// ---------------------------------------------------------------- descriptions
struct option_description {
    std::string long_name;
    std::string short_name;
    std::string help;
    bool takes_value;
    bool has_default;
    std::string default_text;
};

class options_description;

class options_description_easy_init {
    options_description *owner_;
public:
    explicit options_description_easy_init(options_description *o) : owner_(o) {}
    options_description_easy_init &operator()(const char *name, const char *help);
    options_description_easy_init &operator()(const char *name, const value_semantic *s, const char *help);
    options_description_easy_init &operator()(const char *name, const value_semantic *s);
};

class options_description {
    std::string caption_;
    std::vector<option_description> opts_;
    friend class options_description_easy_init;
    void push(const char *name, const value_semantic *s, const char *help) {
        option_description d;
        std::string n(name);
        size_t c = n.find(',');
        d.long_name = n.substr(0, c);
        d.short_name = c == std::string::npos ? "" : n.substr(c + 1);
        d.help = help ? help : "";
        d.takes_value = s != 0;
        d.has_default = s && s->has_default;
        d.default_text = s ? s->default_text : "";
        delete s;
        opts_.push_back(d);
    }
public:
    options_description() {}
    explicit options_description(const std::string &caption, unsigned = 80, unsigned = 40) : caption_(caption) {}
    options_description_easy_init add_options() { return options_description_easy_init(this); }
    options_description &add(const options_description &o) {
        opts_.insert(opts_.end(), o.opts_.begin(), o.opts_.end());
        return *this;
    }
    const std::vector<option_description> &options() const { return opts_; }
    const std::string &caption() const { return caption_; }

    const option_description *find_long(const std::string &k) const {
        const option_description *prefix_hit = 0;
        int n_prefix = 0;
        for (size_t i = 0; i < opts_.size(); i++) {
            if (opts_[i].long_name == k) return &opts_[i];
            if (opts_[i].long_name.compare(0, k.size(), k) == 0) {
                if (!prefix_hit || prefix_hit->long_name != opts_[i].long_name) n_prefix++;
                prefix_hit = &opts_[i];
            }
        }
        if (n_prefix == 1) return prefix_hit;
        if (n_prefix > 1) throw error("option '--" + k + "' is ambiguous");
        return 0;
    }
    const option_description *find_short(const std::string &k) const {
        for (size_t i = 0; i < opts_.size(); i++)
            if (!opts_[i].short_name.empty() && opts_[i].short_name == k) return &opts_[i];
        return 0;
    }
};

inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const char *help) {
    owner_->push(name, 0, help);
    return *this;
}
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const value_semantic *s,
                                                                              const char *help) {
    owner_->push(name, s, help);
    return *this;
}
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const value_semantic *s) {
    owner_->push(name, s, "");
    return *this;
}

inline std::ostream &operator<<(std::ostream &o, const options_description &d) {
    if (!d.caption().empty()) o << d.caption() << ":\n";
    for (size_t i = 0; i < d.options().size(); i++) {
        const option_description &x = d.options()[i];
        o << "  ";
        if (!x.short_name.empty()) o << "-" << x.short_name << " [ --" << x.long_name << " ]";
        else o << "--" << x.long_name;
        if (x.takes_value) o << " arg";
        if (x.has_default) o << " (=" << x.default_text << ")";
        o << "  " << x.help << "\n";
    }
    return o;
}

class positional_options_description {
    std::vector<std::string> names_;
public:
    positional_options_description &add(const char *name, int max_count) {
        for (int i = 0; i < (max_count < 0 ? 1 : max_count); i++) names_.push_back(name);
        return *this;
    }
    const std::vector<std::string> &names() const { return names_; }
};

// ---------------------------------------------------------------- parsing
template <class charT> struct basic_option {
    std::string string_key;
    int position_key;
    std::vector<std::basic_string<charT> > value;
    basic_option() : position_key(-1) {}
};
typedef basic_option<char> option;

template <class charT> struct basic_parsed_options {
    std::vector<basic_option<charT> > options;
    const options_description *description;
    basic_parsed_options() : description(0) {}
    explicit basic_parsed_options(const options_description *d) : description(d) {}
};
typedef basic_parsed_options<char> parsed_options;

class command_line_parser {
    std::vector<std::string> args_;
    const options_description *desc_;
    const positional_options_description *pos_;
public:
    command_line_parser(int argc, const char *const argv[]) : desc_(0), pos_(0) {
        for (int i = 1; i < argc; i++) args_.push_back(argv[i]);
    }
    command_line_parser &options(const options_description &d) { desc_ = &d; return *this; }
    command_line_parser &positional(const positional_options_description &p) { pos_ = &p; return *this; }
    parsed_options run() {
        parsed_options out(desc_);
        size_t n_pos = 0;
        for (size_t i = 0; i < args_.size(); i++) {
            const std::string &a = args_[i];
            const option_description *d = 0;
            std::string inline_value;
            bool has_inline = false;
            if (a.size() > 2 && a.compare(0, 2, "--") == 0) {
                std::string k = a.substr(2);
                size_t eq = k.find('=');
                if (eq != std::string::npos) { inline_value = k.substr(eq + 1); k = k.substr(0, eq); has_inline = true; }
                d = desc_ ? desc_->find_long(k) : 0;
                if (!d) throw error("unrecognised option '" + a + "'");
            } else if (a.size() >= 2 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9') && a[1] != '.') {
                d = desc_ ? desc_->find_short(a.substr(1, 1)) : 0;
                if (!d) throw error("unrecognised option '" + a + "'");
                if (a.size() > 2) { inline_value = a.substr(2); has_inline = true; }
            } else {
                option o;
                if (pos_ && n_pos < pos_->names().size()) o.string_key = pos_->names()[n_pos];
                else throw error("too many positional options have been specified on the command line");
                o.position_key = (int)n_pos++;
                o.value.push_back(a);
                out.options.push_back(o);
                continue;
            }
            option o;
            o.string_key = d->long_name;
            if (d->takes_value) {
                if (has_inline) o.value.push_back(inline_value);
                else if (i + 1 < args_.size()) o.value.push_back(args_[++i]);
                else throw error("the required argument for option '--" + d->long_name + "' is missing");
            }
            out.options.push_back(o);
        }
        return out;
    }
};

inline parsed_options parse_command_line(int argc, const char *const argv[], const options_description &d) {
    return command_line_parser(argc, argv).options(d).run();
}

inline parsed_options parse_config_file(std::istream &is, const options_description &d, bool = false) {
    parsed_options out(&d);
    std::string line;
    while (std::getline(is, line)) {
        size_t h = line.find('#');
        if (h != std::string::npos) line = line.substr(0, h);
        size_t b = line.find_first_not_of(" \t\r\n");
        if (b == std::string::npos) continue;
        size_t e = line.find_last_not_of(" \t\r\n");
        line = line.substr(b, e - b + 1);
        std::string k = line, v;
        size_t eq = line.find('=');
        if (eq != std::string::npos) {
            k = line.substr(0, eq);
            v = line.substr(eq + 1);
            size_t kb = k.find_last_not_of(" \t");
            k = kb == std::string::npos ? "" : k.substr(0, kb + 1);
            size_t vb = v.find_first_not_of(" \t");
            v = vb == std::string::npos ? "" : v.substr(vb);
        }
        const option_description *od = d.find_long(k);
        if (!od) throw error("unrecognised option '" + k + "'");
        option o;
        o.string_key = od->long_name;
        if (od->takes_value) o.value.push_back(v);
        out.options.push_back(o);
    }
    return out;
}

class variables_map : public std::map<std::string, variable_value> {
    typedef std::map<std::string, variable_value> base;
public:
    const variable_value &operator[](const std::string &k) const {
        static const variable_value none;
        base::const_iterator it = base::find(k);
        return it == base::end() ? none : it->second;
    }
    void set(const std::string &k, const std::string &v, bool defaulted = false) {
        base::iterator it = base::find(k);
        if (it == base::end()) base::insert(std::make_pair(k, variable_value(v, defaulted)));
        else if (it->second.defaulted() && !defaulted) it->second = variable_value(v, false);
        // first explicit store wins, as in boost
    }
};

inline void store(const parsed_options &p, variables_map &vm) {
    for (size_t i = 0; i < p.options.size(); i++) {
        const option &o = p.options[i];
        vm.set(o.string_key, o.value.empty() ? std::string("") : o.value[0], false);
    }
    if (p.description)
        for (size_t i = 0; i < p.description->options().size(); i++) {
            const option_description &d = p.description->options()[i];
            if (d.has_default) vm.set(d.long_name, d.default_text, true);
        }
}

inline void notify(variables_map &) {}

}} // namespace boost::program_options
#endif
