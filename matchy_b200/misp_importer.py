"""MISP JSON -> DatabaseBuilder entries: the `-f misp` input of `matchy build` (crates/matchy/src/misp_importer.rs,
bin/commands/build_cmd.rs:311-335).  Off the scan path: it only decides which indicators reach the builder
and which flat metadata map rides on each — the .mxy that comes out is scanned by the same kernels as any other.

One file = one `{"Event": {...}}` document.  Event fields become event_info / event_uuid / threat_level / analysis /
event_date / org_name / tags (misp_importer.rs:816-871); every attribute (direct, or inside an Object, which adds object_type /
object_comment, :778-814) adds type / category / to_ids / comment and its own tags (:873-928); the attribute type picks
add_ip / add_literal and how a composite value is split (:930-1083)."""
import json
import os
import sys

_LITERAL_TYPES = frozenset("""domain hostname md5 sha1 sha224 sha256 sha384 sha512 sha512/224 sha512/256 sha3-224 sha3-256 sha3-384
sha3-512 ssdeep imphash tlsh authentihash vhash cdhash pehash impfuzzy telfhash filename filename-pattern email email-src email-dst
email-reply-to email-subject email-body user-agent http-method mac-address mac-eui-64 AS btc xmr dash yara snort sigma pattern-in-file
pattern-in-traffic pattern-in-memory mutex regkey regkey|value""".split()) | {"named pipe"}
_FILENAME_HASH = frozenset("filename|" + h for h in "md5 sha1 sha256 sha384 sha512 imphash ssdeep tlsh authentihash vhash pehash impfuzzy".split())
_IGNORED = frozenset("comment text other link datetime size-in-bytes counter float hex port attachment malware-sample".split())
_THREAT = {1: "High", 2: "Medium", 3: "Low"}
_ANALYSIS = {0: "Initial", 1: "Ongoing", 2: "Complete"}


class MispError(ValueError):
    pass


def _opt_str(d, key):
    v = d.get(key)
    if v is not None and not isinstance(v, str):
        raise MispError("invalid type: %s is not a string" % key)
    return v


def _u8(v, key):
    """deserialize_u8_flexible (misp_importer.rs:57-80): a number or a string of one, null = absent."""
    if v is None:
        return None
    if isinstance(v, bool) or not isinstance(v, (int, str)):
        raise MispError("expected number or string for u8 (%s)" % key)
    try:
        n = int(v) if isinstance(v, int) or (v.isascii() and v.lstrip("+").isdigit()) else -1
    except ValueError:
        n = -1
    if not 0 <= n <= 255:
        raise MispError("number out of u8 range (%s)" % key)
    return n


def _value_text(v):
    """deserialize_value (misp_importer.rs:40-55): strings as they are, numbers and booleans printed, anything else empty."""
    if isinstance(v, str):
        return v
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, float)):
        return json.dumps(v)
    return ""


def _tag_names(d):
    tags = d.get("Tag") or []
    if not isinstance(tags, list) or any(not isinstance(t, dict) or not isinstance(t.get("name"), str) for t in tags):
        raise MispError("invalid Tag list")
    return [t["name"] for t in tags]


def domain_of_url(url):
    """extract_domain_from_url (misp_importer.rs:1085-1123)."""
    url = url.strip()
    p = url.find("://")
    rest = url[p + 3:] if p >= 0 else url
    end = len(rest)
    for sep in "/?#":  # first '/', else first '?', else first '#'
        k = rest.find(sep)
        if k >= 0:
            end = k
            break
    host = rest[:end]
    c = host.rfind(":")
    if c >= 0 and all(ch.isnumeric() for ch in host[c + 1:]):
        host = host[:c]
    return host or None


def event_metadata(ev):
    md = {}
    if _opt_str(ev, "info") is not None:
        md["event_info"] = ev["info"]
    if _opt_str(ev, "uuid") is not None:
        md["event_uuid"] = ev["uuid"]
    t = _u8(ev.get("threat_level_id"), "threat_level_id")
    if t is not None:
        md["threat_level"] = _THREAT.get(t, "Undefined")
    a = _u8(ev.get("analysis"), "analysis")
    if a is not None:
        md["analysis"] = _ANALYSIS.get(a, "Unknown")
    if _opt_str(ev, "date") is not None:
        md["event_date"] = ev["date"]
    org = ev.get("Orgc")
    if isinstance(org, dict) and isinstance(org.get("name"), str):
        md["org_name"] = org["name"]
    names = _tag_names(ev)
    if names:
        md["tags"] = ",".join(names)
    return md


def add_indicators(b, attr_type, value, md):
    """extract_indicators (misp_importer.rs:930-1083).  Returns how many entries went to the builder."""
    if not value.strip():
        return 0
    bar = value.find("|")
    if attr_type in ("ip-src", "ip-dst", "ip", "ip-src/netmask", "ip-dst/netmask"):
        b.add_ip(value, md)
    elif attr_type in ("ip-src|port", "ip-dst|port"):
        if bar < 0:
            return 0
        b.add_ip(value[:bar], md)
    elif attr_type == "domain|ip":
        if bar < 0:
            return 0
        b.add_literal(value[:bar], md)
        b.add_ip(value[bar + 1:], md)
        return 2
    elif attr_type in ("url", "uri"):
        d = domain_of_url(value)
        if d:
            b.add_literal(d, md)
        b.add_literal(value, md)
        return 2 if d else 1
    elif attr_type in _FILENAME_HASH:
        if bar < 0:
            return 0
        b.add_literal(value[:bar], md)
        b.add_literal(value[bar + 1:], md)
        return 2
    elif attr_type in _LITERAL_TYPES:
        b.add_literal(value, md)
    elif attr_type in _IGNORED:
        return 0
    elif len(value.encode()) < 1000:
        b.add_literal(value, md)
    else:
        return 0
    return 1


def _attribute(b, attr, base):
    if not isinstance(attr, dict) or not isinstance(attr.get("type"), str) or "value" not in attr:
        raise MispError("attribute without type/value")
    md = dict(base)
    md["type"] = attr["type"]
    if _opt_str(attr, "category") is not None:
        md["category"] = attr["category"]
    if attr.get("to_ids") is not None:
        if not isinstance(attr["to_ids"], bool):
            raise MispError("invalid type: to_ids is not a boolean")
        md["to_ids"] = attr["to_ids"]
    if _opt_str(attr, "comment"):
        md["comment"] = attr["comment"]
    names = _tag_names(attr)
    if names:
        md["tags"] = (md["tags"] + "," if md.get("tags") else "") + ",".join(names)
    return add_indicators(b, attr["type"], _value_text(attr["value"]), md)


def add_event(b, ev):
    """process_event (misp_importer.rs:778-814)."""
    if not isinstance(ev, dict):
        raise MispError("Event is not an object")
    base = event_metadata(ev)
    n = 0
    for attr in ev.get("Attribute") or []:
        n += _attribute(b, attr, base)
    for obj in ev.get("Object") or []:
        if not isinstance(obj, dict) or not isinstance(obj.get("name"), str):
            raise MispError("object without a name")
        md = dict(base)
        md["object_type"] = obj["name"]
        if _opt_str(obj, "comment") is not None:
            md["object_comment"] = obj["comment"]
        for attr in obj.get("Attribute") or []:
            n += _attribute(b, attr, md)
    return n


class _Staged:
    def __init__(self): self.entries = []
    def add_ip(self, key, md): self.entries.append((True, key, md))
    def add_literal(self, key, md): self.entries.append((False, key, md))


def add_misp_files(b, paths, warn=None):
    """MispImporter::build_from_files (misp_importer.rs:300-373): one event per file, streamed into the builder; files that are
    not MISP events are skipped with a warning, files that look like one but do not parse are errors."""
    b.set_database_type("MISP-ThreatIntel")
    b.set_description("en", "Threat intelligence database from MISP JSON feeds")
    warn = warn or sys.stderr
    skipped, events, total = [], 0, 0
    for path in paths:
        try:
            with open(path, encoding="utf-8") as f:
                text = f.read()
        except (OSError, UnicodeDecodeError) as e:
            raise MispError("Failed to read file: %s" % e)
        name = os.path.basename(os.fspath(path)) or "unknown"
        try:
            doc = json.loads(text)
            if not isinstance(doc, dict) or "Event" not in doc:
                raise MispError("missing field `Event`")
            staged = _Staged()  # (serde parses the whole document before anything reaches the builder)
            add_event(staged, doc["Event"])
            events += 1
        except (json.JSONDecodeError, MispError) as e:
            if name in ("manifest.json", "hashes.csv"):
                skipped.append((name, "metadata file"))
            elif text.lstrip().startswith("{") and '"Event"' in text:
                raise MispError("Failed to parse MISP JSON in %s: %s" % (name, e))
            else:
                skipped.append((name, "not a MISP event"))
            continue
        for ip, key, md in staged.entries:
            (b.add_ip if ip else b.add_literal)(key, md)
        total += len(staged.entries)
    if skipped:
        print("Warning: Skipped %d non-MISP file(s):" % len(skipped), file=warn)
        for name, why in skipped:
            print("  - %s: %s" % (name, why), file=warn)
    if events == 0:
        raise MispError("No valid MISP events found in provided files")
    return total
