// xml_lite — a small DOM reader for the reference's input decks.
//
// The reference parses `<dir>/input.xml` with pugixml (src/simulator/setup.cpp:32-35).
// Its decks are XML *fragments* (several top-level elements, no single root) using
// only elements, attributes, comments and an XML declaration.  This reader covers
// that subset with pugixml's observable semantics for the calls setup.cpp makes:
//   child(name) / children(name) / children()   document order, first match
//   attribute(name)  -> truthiness, value() (empty string when absent)
//   as_double()      -> strtod, 0.0 when absent      as_int() -> strtol, 0 when absent
#ifndef MCB_XML_LITE_H
#define MCB_XML_LITE_H

#include <string>
#include <utility>
#include <vector>

namespace mcb {

struct XmlAttr {
    bool present = false;
    std::string text;
    explicit operator bool() const { return present; }
    const std::string& value() const { return text; }
    double as_double() const;
    int as_int() const;
};

struct XmlNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<XmlNode> kids;

    XmlAttr attribute(const std::string& n) const;
    // first child with that name, or nullptr
    const XmlNode* child(const std::string& n) const;
    std::vector<const XmlNode*> children(const std::string& n) const;
    std::vector<const XmlNode*> children() const;
};

// Parses a whole file into a synthetic root whose kids are the top-level elements.
// Returns false and sets err on malformed input or unreadable file.
bool xml_parse_file(const std::string& path, XmlNode& root, std::string& err);
bool xml_parse_string(const std::string& text, XmlNode& root, std::string& err);

}  // namespace mcb
#endif
