"""Mirror of the reference's `matchy_extractor::Extractor` (crates/matchy-extractor/src/lib.rs:17-133, 345-488),
running on the device: `extract_from_chunk` == tokenize + validate kernels."""
import ipaddress

from . import engine as E


class ExtractorError(Exception):
    pass


class Match:
    """`Match { item, span }` — item_type is the reference's type name ("Domain", "IPv4", "MD5", …)."""
    __slots__ = ("item_type", "span", "_chunk")

    def __init__(self, item_type, span, chunk):
        self.item_type, self.span, self._chunk = item_type, span, chunk

    def type_name(self):
        return self.item_type

    def as_bytes(self):
        return bytes(self._chunk[self.span[0]:self.span[1]])

    def as_str(self, _input=None):
        return self.as_bytes().decode("utf-8")

    def as_value(self):
        """`ExtractedItem::as_value`: canonical text for IP addresses (lib.rs:300-311), raw text otherwise."""
        s = self.as_str()
        if self.item_type == "IPv4":
            return str(ipaddress.IPv4Address(s))
        if self.item_type == "IPv6":
            return _rust_ipv6_display(ipaddress.IPv6Address(s))
        return s

    def __repr__(self):
        return "Match(%s, %r, span=%r)" % (self.item_type, self.as_bytes(), self.span)


def _rust_ipv6_display(a):
    if a.ipv4_mapped is not None:
        return "::ffff:" + str(a.ipv4_mapped)
    return a.compressed


class ExtractorBuilder:
    def __init__(self):
        self._f = {"domains": True, "emails": True, "ipv4": True, "ipv6": True, "hashes": True,
                   "bitcoin": True, "ethereum": True, "monero": True}
        self._min_labels = 2
        self._boundaries = True

    def extract_domains(self, enable): self._f["domains"] = bool(enable); return self
    def extract_emails(self, enable): self._f["emails"] = bool(enable); return self
    def extract_ipv4(self, enable): self._f["ipv4"] = bool(enable); return self
    def extract_ipv6(self, enable): self._f["ipv6"] = bool(enable); return self
    def extract_hashes(self, enable): self._f["hashes"] = bool(enable); return self
    def extract_bitcoin(self, enable): self._f["bitcoin"] = bool(enable); return self
    def extract_ethereum(self, enable): self._f["ethereum"] = bool(enable); return self
    def extract_monero(self, enable): self._f["monero"] = bool(enable); return self
    def min_domain_labels(self, n): self._min_labels = int(n); return self
    def require_word_boundaries(self, r): self._boundaries = bool(r); return self

    def build(self, engine=None, device=0):
        if self._min_labels != 2 or not self._boundaries:
            raise ExtractorError("the device extractor implements the defaults only (min_domain_labels=2, require_word_boundaries=true)")
        return Extractor(dict(self._f), engine, device)


class Extractor:
    """`Extractor::new()` enables everything, crypto-address extractors (bitcoin/ethereum/monero) included — like the
    reference.  `.extract_bitcoin(False).extract_ethereum(False).extract_monero(False)` == `--extractors=-crypto`."""

    def __init__(self, enabled=None, engine=None, device=0):
        self._f = enabled or {"domains": True, "emails": True, "ipv4": True, "ipv6": True, "hashes": True,
                              "bitcoin": True, "ethereum": True, "monero": True}
        self._engine = engine
        self._device = device

    @staticmethod
    def new(engine=None, device=0):
        return Extractor(None, engine, device)

    @staticmethod
    def builder():
        return ExtractorBuilder()

    def extract_domains(self): return self._f["domains"]
    def extract_emails(self): return self._f["emails"]
    def extract_ipv4(self): return self._f["ipv4"]
    def extract_ipv6(self): return self._f["ipv6"]
    def extract_hashes(self): return self._f["hashes"]
    def extract_bitcoin(self): return self._f["bitcoin"]
    def extract_ethereum(self): return self._f["ethereum"]
    def extract_monero(self): return self._f["monero"]
    def min_domain_labels(self): return 2

    def flags(self):
        f = self._f
        return ((E.X_DOMAINS if f["domains"] else 0) | (E.X_EMAILS if f["emails"] else 0) | (E.X_IPV4 if f["ipv4"] else 0) |
                (E.X_IPV6 if f["ipv6"] else 0) | (E.X_HASHES if f["hashes"] else 0) | (E.X_BITCOIN if f["bitcoin"] else 0) |
                (E.X_ETHEREUM if f["ethereum"] else 0) | (E.X_MONERO if f["monero"] else 0))

    def device_flags(self):
        fl = self.flags()
        if fl & ~E.X_SUPPORTED:
            raise ExtractorError("unknown extractor flags")
        return fl

    def bind(self, engine):
        self._engine = engine
        return self

    def _eng(self):
        if self._engine is None:
            self._engine = E.Engine(self._device)
        return self._engine

    # reference order of extract_from_chunk: IPv6, IPv4, e-mail, domain, hashes, bitcoin, ethereum, monero; ascending offset inside a type
    _ORDER = {3: 0, 2: 1, 1: 2, 0: 3, 4: 4, 5: 4, 6: 4, 7: 4, 8: 4, 9: 5, 10: 6, 11: 7}

    def extract_from_chunk(self, chunk):
        items = self._eng().extract(chunk, self.device_flags())
        items.sort(key=lambda t: (self._ORDER[t[0]], t[1]))
        return [Match(E.ITEM_TYPE_NAMES[t], (s, e), chunk) for t, s, e in items]

    # extract_from_line order (lib.rs:1472-1521): domains, IPv4, e-mails, IPv6, hashes, bitcoin, ethereum, monero
    _LINE_ORDER = {0: 0, 2: 1, 1: 2, 3: 3, 4: 4, 5: 4, 6: 4, 7: 4, 8: 4, 9: 5, 10: 6, 11: 7}

    def extract_from_line(self, line):
        items = self._eng().extract(line, self.device_flags())
        items.sort(key=lambda t: (self._LINE_ORDER[t[0]], t[1]))
        return [Match(E.ITEM_TYPE_NAMES[t], (s, e), line) for t, s, e in items]
