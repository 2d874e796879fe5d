// config.h -- INI reader/writer with the interface of the reference's tiny::config (libs/headeronly/tinyconfig.hpp:37-57):
// "[section]" headers, "key = value" settings, and repeated "<key = value, ...>" vector entries (single line or
// one pair per line between "<" and ">").  conf.ini files written for AeroFLEX load unchanged.  Independent code.
#pragma once
#include <algorithm>
#include <cctype>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

namespace tiny {

class config {
public:
    using entry = std::unordered_map<std::string, std::string>;
    std::unordered_map<std::string, entry> config;  // same public member names as the reference (no user constructors)
    std::unordered_map<std::string, std::vector<entry>> config_vec;
    std::vector<std::string> sections;

    bool read(const std::string& path) {
        std::ifstream f(path);
        if (!f.is_open()) throw std::runtime_error("Config file: " + path + " not found");
        std::string line, section = "_";
        int line_nb = 0;
        while (std::getline(f, line)) {
            ++line_nb;
            const auto lb = line.find('[');
            const auto la = line.find('<');
            if (lb != std::string::npos) {
                const auto rb = line.find(']');
                if (rb == std::string::npos) throw std::runtime_error("Closing character is missing on line: " + std::to_string(line_nb));
                section = strip(line.substr(lb + 1, rb - lb - 1));
                config[section];
                sections.push_back(section);
            } else if (la != std::string::npos) {
                entry e;
                const auto last = line.find_last_not_of(" \t\f\v\n\r");
                if (la == last) {  // block form: one "k = v," per line until ">"
                    while (std::getline(f, line) && line.find('>') == std::string::npos) {
                        ++line_nb;
                        pair(e, line, line.find(','), line_nb, true);
                    }
                } else if (line.find('>') != std::string::npos) {  // inline form
                    std::string rest = line.substr(la + 1);
                    for (auto c = rest.find(','); c != std::string::npos; c = rest.find(',')) {
                        pair(e, rest, c, line_nb, true);
                        rest = rest.substr(c + 1);
                    }
                    pair(e, rest, rest.find('>'), line_nb, true);
                }
                config_vec[section].push_back(std::move(e));
            } else {
                pair(config[section], line, line.length(), line_nb, false);
            }
        }
        return !f.bad();
    }

    bool write(const std::string& path) {
        std::ofstream f(path);
        if (!f.is_open()) return false;
        for (const auto& s : sections) {
            f << "[" << s << "] \n";
            for (const auto& [k, v] : config[s]) f << k << " = " << v << "\n";
            for (const auto& e : config_vec[s]) {
                if (e.size() > 1) {
                    f << "<\n";
                    for (const auto& [k, v] : e) f << "\t" << k << " = " << v << ",\n";
                    f << ">\n";
                } else if (e.size() == 1) f << "<" << e.begin()->first << " = " << e.begin()->second << ">\n";
            }
            f << "\n";
        }
        return !f.bad();
    }

    template <typename T> T get(const std::string& section, const std::string& key) { return convert<T>(find(find_section(config, section), key)); }
    template <typename T> T get_i(const std::string& section, const std::string& key, int idx) {
        return convert<T>(find(find_section(config_vec, section).at((size_t)idx), key));
    }
    bool has(const std::string& section, const std::string& key) { return find_section(config, section).count(key) != 0; }
    bool has_i(const std::string& section, const std::string& key, int idx) { return find_section(config_vec, section).at((size_t)idx).count(key) != 0; }
    int how_many(const std::string& section) { return (int)find_section(config_vec, section).size(); }

private:
    static std::string strip(std::string s) {
        s.erase(std::remove_if(s.begin(), s.end(), [](unsigned char c) { return std::isspace(c); }), s.end());
        return s;
    }
    static void pair(entry& e, const std::string& line, std::size_t end, int line_nb, bool need_end) {
        const auto eq = line.find('=');
        if (eq == std::string::npos) return;
        if (end == std::string::npos) {
            if (need_end) throw std::runtime_error("Closing character is missing on line: " + std::to_string(line_nb));
            end = line.length();
        }
        e[strip(line.substr(0, eq))] = strip(line.substr(eq + 1, end - eq - 1));
    }
    template <class M> static typename M::mapped_type& find_section(M& m, const std::string& s) {
        auto it = m.find(s);
        if (it == m.end()) throw std::runtime_error("Section [" + s + "] not found");
        return it->second;
    }
    static const std::string& find(const entry& e, const std::string& key) {
        auto it = e.find(key);
        if (it == e.end()) throw std::runtime_error("Setting (" + key + ") not found");
        return it->second;
    }
    template <typename T> static T convert(std::string s) {
        if constexpr (std::is_same<T, std::string>::value) return s;
        else if constexpr (std::is_same<T, bool>::value) {
            std::transform(s.begin(), s.end(), s.begin(), ::tolower);
            if (s == "1" || s == "true" || s == "yes" || s == "on") return true;
            if (s == "0" || s == "false" || s == "no" || s == "off") return false;
            throw std::runtime_error("'" + s + "' is not a valid boolean value.");
        } else {
            T v{};
            std::istringstream(s) >> v;
            return v;
        }
    }
};

}  // namespace tiny
