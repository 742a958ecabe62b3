#include "xml_lite.h"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace mcb {

double XmlAttr::as_double() const { return present ? std::strtod(text.c_str(), nullptr) : 0.0; }
int XmlAttr::as_int() const { return present ? (int)std::strtol(text.c_str(), nullptr, 10) : 0; }

XmlAttr XmlNode::attribute(const std::string& n) const
{
    XmlAttr a;
    for (const auto& kv : attrs) {
        if (kv.first == n) { a.present = true; a.text = kv.second; break; }
    }
    return a;
}
const XmlNode* XmlNode::child(const std::string& n) const
{
    for (const auto& k : kids) { if (k.name == n) return &k; }
    return nullptr;
}
std::vector<const XmlNode*> XmlNode::children(const std::string& n) const
{
    std::vector<const XmlNode*> v;
    for (const auto& k : kids) { if (k.name == n) v.push_back(&k); }
    return v;
}
std::vector<const XmlNode*> XmlNode::children() const
{
    std::vector<const XmlNode*> v;
    for (const auto& k : kids) v.push_back(&k);
    return v;
}

namespace {

struct Parser {
    const std::string& s;
    size_t i = 0;
    std::string err;
    explicit Parser(const std::string& text) : s(text) {}

    bool starts(const char* lit) const { return s.compare(i, std::char_traits<char>::length(lit), lit) == 0; }
    void skip_ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) i++; }
    bool fail(const std::string& m)
    {
        size_t line = 1;
        for (size_t k = 0; k < i && k < s.size(); k++) if (s[k] == '\n') line++;
        std::ostringstream o; o << m << " (line " << line << ")"; err = o.str();
        return false;
    }
    static bool name_char(char c) { return std::isalnum((unsigned char)c) || c == '_' || c == '-' || c == '.' || c == ':'; }

    static std::string decode(const std::string& v)
    {
        if (v.find('&') == std::string::npos) return v;
        std::string o;
        for (size_t k = 0; k < v.size();) {
            if (v[k] == '&') {
                static const char* ent[] = {"&amp;", "&lt;", "&gt;", "&quot;", "&apos;"};
                static const char rep[] = {'&', '<', '>', '"', '\''};
                bool hit = false;
                for (int e = 0; e < 5; e++) {
                    const size_t L = std::char_traits<char>::length(ent[e]);
                    if (v.compare(k, L, ent[e]) == 0) { o += rep[e]; k += L; hit = true; break; }
                }
                if (hit) continue;
            }
            o += v[k++];
        }
        return o;
    }

    // skips comments, processing instructions, doctype and text; stops at '<' of an element/closing tag or EOF
    bool skip_misc()
    {
        for (;;) {
            while (i < s.size() && s[i] != '<') i++;
            if (i >= s.size()) return true;
            if (starts("<!--")) {
                const size_t e = s.find("-->", i + 4);
                if (e == std::string::npos) return fail("unterminated comment");
                i = e + 3;
            } else if (starts("<?")) {
                const size_t e = s.find("?>", i + 2);
                if (e == std::string::npos) return fail("unterminated processing instruction");
                i = e + 2;
            } else if (starts("<![CDATA[")) {
                const size_t e = s.find("]]>", i + 9);
                if (e == std::string::npos) return fail("unterminated CDATA");
                i = e + 3;
            } else if (starts("<!")) {
                const size_t e = s.find('>', i + 2);
                if (e == std::string::npos) return fail("unterminated declaration");
                i = e + 1;
            } else {
                return true;
            }
        }
    }

    bool parse_element(XmlNode& out)
    {
        i++;  // '<'
        size_t b = i;
        while (i < s.size() && name_char(s[i])) i++;
        if (i == b) return fail("expected element name");
        out.name = s.substr(b, i - b);
        for (;;) {
            skip_ws();
            if (i >= s.size()) return fail("unterminated start tag <" + out.name);
            if (s[i] == '/') {
                if (i + 1 < s.size() && s[i + 1] == '>') { i += 2; return true; }
                return fail("stray '/' in tag <" + out.name);
            }
            if (s[i] == '>') { i++; break; }
            b = i;
            while (i < s.size() && name_char(s[i])) i++;
            if (i == b) return fail("bad attribute in <" + out.name);
            std::string an = s.substr(b, i - b);
            skip_ws();
            if (i >= s.size() || s[i] != '=') return fail("attribute '" + an + "' without value");
            i++;
            skip_ws();
            if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) return fail("attribute '" + an + "' value not quoted");
            const char q = s[i++];
            b = i;
            while (i < s.size() && s[i] != q) i++;
            if (i >= s.size()) return fail("unterminated attribute value");
            out.attrs.emplace_back(an, decode(s.substr(b, i - b)));
            i++;
        }
        // content
        for (;;) {
            if (!skip_misc()) return false;
            if (i >= s.size()) return fail("missing </" + out.name + ">");
            if (starts("</")) {
                i += 2;
                b = i;
                while (i < s.size() && name_char(s[i])) i++;
                if (s.substr(b, i - b) != out.name) return fail("mismatched </" + s.substr(b, i - b) + ">, open <" + out.name + ">");
                skip_ws();
                if (i >= s.size() || s[i] != '>') return fail("bad closing tag");
                i++;
                return true;
            }
            out.kids.emplace_back();
            if (!parse_element(out.kids.back())) return false;
        }
    }

    bool parse_document(XmlNode& root)
    {
        root = XmlNode();
        for (;;) {
            if (!skip_misc()) return false;
            if (i >= s.size()) return true;
            if (starts("</")) return fail("closing tag without an open element");
            root.kids.emplace_back();
            if (!parse_element(root.kids.back())) return false;
        }
    }
};

}  // namespace

bool xml_parse_string(const std::string& text, XmlNode& root, std::string& err)
{
    Parser p(text);
    if (!p.parse_document(root)) { err = p.err; return false; }
    return true;
}

bool xml_parse_file(const std::string& path, XmlNode& root, std::string& err)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "cannot open " + path; return false; }
    std::ostringstream ss;
    ss << f.rdbuf();
    return xml_parse_string(ss.str(), root, err);
}

}  // namespace mcb
