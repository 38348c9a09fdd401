#include "CdlIO.hpp"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace ddc_host {
namespace {

struct Tok {
    enum Kind { WORD, SYM, STR, END } kind;
    std::string text;
};

class Lexer {
public:
    explicit Lexer(const std::string& s)
        : _s(s)
    {
    }
    Tok next()
    {
        skip();
        if (_i >= _s.size())
            return { Tok::END, "" };
        char c = _s[_i];
        if (c == '"') {
            size_t j = _i + 1;
            std::string out;
            while (j < _s.size() && _s[j] != '"') {
                if (_s[j] == '\\' && j + 1 < _s.size())
                    j++;
                out += _s[j++];
            }
            _i = j + 1;
            return { Tok::STR, out };
        }
        if (std::string("{}(),;=:").find(c) != std::string::npos) {
            _i++;
            return { Tok::SYM, std::string(1, c) };
        }
        size_t j = _i;
        while (j < _s.size() && !std::isspace((unsigned char)_s[j]) && std::string("{}(),;=:\"").find(_s[j]) == std::string::npos)
            j++;
        Tok t { Tok::WORD, _s.substr(_i, j - _i) };
        _i = j;
        return t;
    }
    Tok peek()
    {
        size_t save = _i;
        Tok t = next();
        _i = save;
        return t;
    }

private:
    void skip()
    {
        for (;;) {
            while (_i < _s.size() && std::isspace((unsigned char)_s[_i]))
                _i++;
            if (_i + 1 < _s.size() && _s[_i] == '/' && _s[_i + 1] == '/') {
                while (_i < _s.size() && _s[_i] != '\n')
                    _i++;
                continue;
            }
            break;
        }
    }
    const std::string& _s;
    size_t _i = 0;
};

[[noreturn]] void bad(const std::string& what) { throw std::runtime_error("ERROR: CDL: " + what); }

void skip_statement(Lexer& lx)
{
    for (;;) {
        Tok t = lx.next();
        if (t.kind == Tok::END || (t.kind == Tok::SYM && t.text == ";"))
            return;
    }
}

double parse_number(const std::string& w)
{
    if (w == "_" || w == "NaN" || w == "NaNf")
        return 0.0;
    char* end = nullptr;
    double v = std::strtod(w.c_str(), &end);
    if (end == w.c_str())
        bad("cannot parse value '" + w + "'");
    return v; // trailing type suffixes (f, L, s, b, u...) are ignored
}

bool is_section(const std::string& w) { return w == "dimensions" || w == "variables" || w == "data" || w == "group" || w == "types"; }

// parses the body of a group up to (and including) its closing brace
void parse_group(Lexer& lx, CdlGroup& g)
{
    std::string section;
    for (;;) {
        Tok t = lx.next();
        if (t.kind == Tok::END)
            bad("unexpected end of file");
        if (t.kind == Tok::SYM && t.text == "}")
            return;
        if (t.kind == Tok::WORD && is_section(t.text) && lx.peek().kind == Tok::SYM && lx.peek().text == ":") {
            lx.next(); // ':'
            if (t.text == "group") {
                Tok name = lx.next();
                Tok brace = lx.next();
                if (name.kind != Tok::WORD || brace.text != "{")
                    bad("malformed group header");
                parse_group(lx, g.groups[name.text]);
            } else
                section = t.text;
            continue;
        }
        if (section == "dimensions") {
            // name = N ;   |   name = UNLIMITED ;   (several may share a line: a = 1, b = 2 ;)
            std::string name = t.text;
            for (;;) {
                Tok eq = lx.next();
                Tok val = lx.next();
                if (eq.text != "=" || val.kind != Tok::WORD)
                    bad("malformed dimension '" + name + "'");
                g.dims[name] = (val.text == "UNLIMITED" || val.text == "unlimited") ? 0 : std::atol(val.text.c_str());
                Tok sep = lx.next();
                if (sep.text == ";")
                    break;
                if (sep.text != ",")
                    bad("malformed dimension list");
                name = lx.next().text;
            }
        } else if (section == "variables") {
            if (t.kind == Tok::SYM && t.text == ":") { // ":global_att = ... ;"
                skip_statement(lx);
                continue;
            }
            Tok nx = lx.peek();
            if (nx.kind == Tok::SYM && nx.text == ":") { // "var:att = ... ;"
                skip_statement(lx);
                continue;
            }
            // type name(d1, d2), name2(d) ;
            std::string type = t.text;
            for (;;) {
                Tok name = lx.next();
                if (name.kind != Tok::WORD)
                    bad("malformed variable declaration");
                CdlVar v;
                v.type = type;
                Tok p = lx.next();
                if (p.text == "(") {
                    for (;;) {
                        Tok d = lx.next();
                        if (d.kind != Tok::WORD)
                            bad("malformed dimension list of '" + name.text + "'");
                        v.dims.push_back(d.text);
                        Tok s = lx.next();
                        if (s.text == ")")
                            break;
                        if (s.text != ",")
                            bad("malformed dimension list of '" + name.text + "'");
                    }
                    p = lx.next();
                }
                g.vars[name.text] = v;
                if (p.text == ";")
                    break;
                if (p.text != ",")
                    bad("malformed variable declaration");
            }
        } else if (section == "data") {
            // name = v, v, ... ;
            Tok eq = lx.next();
            if (eq.text != "=")
                bad("malformed data statement for '" + t.text + "'");
            CdlVar& v = g.vars[t.text];
            v.has_data = true;
            for (;;) {
                Tok val = lx.next();
                if (val.kind == Tok::END)
                    bad("unterminated data statement");
                if (val.kind == Tok::SYM && val.text == ";")
                    break;
                if (val.kind == Tok::SYM)
                    continue; // ',' '{' '}' of vlen / compound notation are ignored
                if (val.kind == Tok::WORD)
                    v.data.push_back(parse_number(val.text));
            }
        } else {
            // global attributes before any section keyword, `types:` bodies, ...
            skip_statement(lx);
        }
    }
}

} // namespace

CdlFile read_cdl(const std::string& path)
{
    std::ifstream in(path);
    if (!in)
        throw std::runtime_error("ERROR: cannot open grid file '" + path + "'");
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string text = ss.str();
    Lexer lx(text);
    Tok kw = lx.next();
    Tok name = lx.next();
    Tok brace = lx.next();
    if (kw.text != "netcdf" || name.kind != Tok::WORD || brace.text != "{")
        bad("'" + path + "' is not CDL text (expected `netcdf <name> {`); build with DDC_HAVE_NETCDF to read binary netCDF");
    CdlFile f;
    f.name = name.text;
    parse_group(lx, f.root);
    return f;
}

std::string format_values(const std::string& first_prefix, const int* v, size_t n, const std::string& cont_indent)
{
    // ncdump keeps lines within 80 columns and continues with an indent
    std::string out = first_prefix;
    size_t col = first_prefix.size();
    size_t nl = first_prefix.rfind('\n');
    if (nl != std::string::npos)
        col = first_prefix.size() - nl - 1;
    for (size_t i = 0; i < n; i++) {
        std::string tok = std::to_string(v[i]);
        tok += (i + 1 < n) ? "," : " ;";
        if (col + tok.size() > 79 && col > cont_indent.size()) {
            out += "\n" + cont_indent;
            col = cont_indent.size();
        }
        out += tok;
        col += tok.size();
        if (i + 1 < n) {
            out += " ";
            col += 1;
        }
    }
    return out;
}

std::string netcdf_name_of(const std::string& filename)
{
    size_t slash = filename.find_last_of('/');
    std::string base = slash == std::string::npos ? filename : filename.substr(slash + 1);
    size_t dot = base.find_last_of('.');
    return dot == std::string::npos ? base : base.substr(0, dot);
}

std::string cdl_path_of(const std::string& filename)
{
    if (filename.size() >= 4 && filename.compare(filename.size() - 4, 4, ".cdl") == 0)
        return filename;
    if (filename.size() >= 3 && filename.compare(filename.size() - 3, 3, ".nc") == 0)
        return filename.substr(0, filename.size() - 3) + ".cdl";
    return filename + ".cdl";
}

} // namespace ddc_host
