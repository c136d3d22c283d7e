// Minimal stand-in for <boost/program_options.hpp>.
//
// Boost is not installed in the build image, but the reference drivers
// (example/gbfs.cu:13, gsssp.cu:12, gpr.cu:13, gtc.cu:13) and the reference
// graphblas/util.hpp:16,39-132 include it and use a small slice of its API:
//   options_description, add_options()(name, value<T>()->default_value(v), help),
//   variables_map, parse_command_line, store, notify, vm["x"].as<T>(), vm.count().
// This header implements exactly that slice so those files compile unchanged.
// It is not derived from Boost sources.
#ifndef GRAPHBLAST_B200_SHIM_BOOST_PROGRAM_OPTIONS_HPP_
#define GRAPHBLAST_B200_SHIM_BOOST_PROGRAM_OPTIONS_HPP_

#include <cassert>
#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#ifndef BOOST_ASSERT
#define BOOST_ASSERT(expr) assert(expr)
#endif

namespace boost {
namespace program_options {

// Type-erased option value -------------------------------------------------
class holder_base {
 public:
  virtual ~holder_base() {}
  virtual holder_base* clone() const = 0;
  virtual bool parse(const std::string& text) = 0;
  virtual const std::type_info& type() const = 0;
  virtual void print(std::ostream& os) const = 0;
};

template <typename T>
struct text_codec {
  static bool decode(const std::string& s, T* out) {
    std::istringstream is(s);
    is >> *out;
    return !is.fail();
  }
};

template <>
struct text_codec<std::string> {
  static bool decode(const std::string& s, std::string* out) {
    *out = s;
    return true;
  }
};

template <>
struct text_codec<bool> {
  static bool decode(const std::string& s, bool* out) {
    if (s == "1" || s == "true" || s == "on" || s == "yes") {
      *out = true;
      return true;
    }
    if (s == "0" || s == "false" || s == "off" || s == "no") {
      *out = false;
      return true;
    }
    return false;
  }
};

template <typename T>
class holder : public holder_base {
 public:
  holder() : v_() {}
  explicit holder(const T& v) : v_(v) {}
  holder_base* clone() const { return new holder<T>(v_); }
  bool parse(const std::string& text) { return text_codec<T>::decode(text, &v_); }
  const std::type_info& type() const { return typeid(T); }
  void print(std::ostream& os) const { os << v_; }
  T v_;
};

// What po::value<T>() returns; ->default_value(v) chains on the pointer.
class value_semantic {
 public:
  virtual ~value_semantic() {}
  virtual holder_base* make_default() const = 0;   // NULL if no default
  virtual holder_base* make_empty() const = 0;
};

template <typename T>
class typed_value : public value_semantic {
 public:
  typed_value() : has_default_(false), def_() {}
  template <typename U>
  typed_value* default_value(const U& v) {
    has_default_ = true;
    def_ = static_cast<T>(v);
    return this;
  }
  typed_value* default_value(const char* v) {
    has_default_ = true;
    text_codec<T>::decode(v, &def_);
    return this;
  }
  holder_base* make_default() const {
    return has_default_ ? new holder<T>(def_) : NULL;
  }
  holder_base* make_empty() const { return new holder<T>(); }

 private:
  bool has_default_;
  T def_;
};

template <typename T>
typed_value<T>* value() {
  return new typed_value<T>();
}

class variable_value {
 public:
  variable_value() {}
  explicit variable_value(holder_base* h) : h_(h) {}
  template <typename T>
  const T& as() const {
    const holder<T>* p = dynamic_cast<const holder<T>*>(h_.get());
    if (p == NULL) throw std::runtime_error("program_options: bad as<T>() cast");
    return p->v_;
  }
  bool empty() const { return !h_; }

 private:
  std::shared_ptr<holder_base> h_;
};

class variables_map : public std::map<std::string, variable_value> {
 public:
  const variable_value& operator[](const std::string& name) const {
    static variable_value none;
    const_iterator it = find(name);
    return it == end() ? none : it->second;
  }
  void insert_value(const std::string& name, holder_base* h) {
    std::map<std::string, variable_value>::operator[](name) = variable_value(h);
  }
};

struct option_spec {
  std::string name;
  std::string help;
  std::shared_ptr<value_semantic> sem;  // empty => flag without argument
};

class options_description;

class options_description_easy_init {
 public:
  explicit options_description_easy_init(options_description* owner)
      : owner_(owner) {}
  options_description_easy_init& operator()(const char* name,
                                            const char* help);
  options_description_easy_init& operator()(const char* name,
                                            value_semantic* sem,
                                            const char* help = "");

 private:
  options_description* owner_;
};

class options_description {
 public:
  options_description() {}
  explicit options_description(const std::string& caption)
      : caption_(caption) {}
  options_description_easy_init add_options() {
    return options_description_easy_init(this);
  }
  const option_spec* find(const std::string& name) const {
    for (size_t i = 0; i < specs_.size(); ++i)
      if (specs_[i].name == name) return &specs_[i];
    return NULL;
  }
  std::string caption_;
  std::vector<option_spec> specs_;
};

inline options_description_easy_init& options_description_easy_init::operator()(
    const char* name, const char* help) {
  option_spec s;
  s.name = name;
  s.help = help;
  owner_->specs_.push_back(s);
  return *this;
}

inline options_description_easy_init& options_description_easy_init::operator()(
    const char* name, value_semantic* sem, const char* help) {
  option_spec s;
  s.name = name;
  s.help = help;
  s.sem.reset(sem);
  owner_->specs_.push_back(s);
  return *this;
}

inline std::ostream& operator<<(std::ostream& os,
                                const options_description& d) {
  os << d.caption_ << ":\n";
  for (size_t i = 0; i < d.specs_.size(); ++i) {
    os << "  --" << d.specs_[i].name;
    if (d.specs_[i].sem) {
      os << " arg";
      std::unique_ptr<holder_base> def(d.specs_[i].sem->make_default());
      if (def) {
        os << " (=";
        def->print(os);
        os << ")";
      }
    }
    os << "  " << d.specs_[i].help << "\n";
  }
  return os;
}

struct parsed_option {
  std::string name;
  std::string text;
  bool has_text;
};

struct parsed_options {
  const options_description* desc;
  std::vector<parsed_option> options;
};

// Tokens that do not start with "--" and are not consumed as an option's
// argument are positional (the .mtx path); like Boost without a
// positional_options_description they are carried but never stored.
inline parsed_options parse_command_line(int argc, const char* const* argv,
                                         const options_description& desc) {
  parsed_options out;
  out.desc = &desc;
  for (int i = 1; i < argc; ++i) {
    std::string tok = argv[i];
    if (tok.size() < 3 || tok[0] != '-' || tok[1] != '-') continue;
    std::string name = tok.substr(2);
    std::string text;
    bool has_text = false;
    size_t eq = name.find('=');
    if (eq != std::string::npos) {
      text = name.substr(eq + 1);
      name = name.substr(0, eq);
      has_text = true;
    }
    const option_spec* spec = desc.find(name);
    if (spec == NULL)
      throw std::runtime_error("unrecognised option '--" + name + "'");
    if (spec->sem && !has_text) {
      if (i + 1 >= argc)
        throw std::runtime_error("option '--" + name + "' needs an argument");
      text = argv[++i];
      has_text = true;
    }
    parsed_option po;
    po.name = name;
    po.text = text;
    po.has_text = has_text;
    out.options.push_back(po);
  }
  return out;
}

inline void store(const parsed_options& parsed, variables_map& vm) {
  const options_description& desc = *parsed.desc;
  for (size_t i = 0; i < parsed.options.size(); ++i) {
    const parsed_option& po = parsed.options[i];
    const option_spec* spec = desc.find(po.name);
    if (vm.count(po.name)) continue;  // first occurrence wins
    if (spec->sem) {
      holder_base* h = spec->sem->make_empty();
      if (!h->parse(po.text)) {
        delete h;
        throw std::runtime_error("bad value '" + po.text + "' for option '--" +
                                 po.name + "'");
      }
      vm.insert_value(po.name, h);
    } else {
      vm.insert_value(po.name, new holder<bool>(true));
    }
  }
  for (size_t i = 0; i < desc.specs_.size(); ++i) {
    const option_spec& s = desc.specs_[i];
    if (!s.sem || vm.count(s.name)) continue;
    holder_base* h = s.sem->make_default();
    if (h != NULL) vm.insert_value(s.name, h);
  }
}

inline void notify(variables_map&) {}

}  // namespace program_options
}  // namespace boost

#endif  // GRAPHBLAST_B200_SHIM_BOOST_PROGRAM_OPTIONS_HPP_
