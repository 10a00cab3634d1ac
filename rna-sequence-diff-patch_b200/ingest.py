"""ingest.py — sequence files -> packed database.

Mirrors the normalisation of the reference's importers (fa_import.py:61-62, import_xml.py:12-13:
T -> U, X -> N); replaces the MongoDB collection scan of IRMethods.search_collection (IR:469) by a packed,
device-resident database.  Like fa_import.py:49 every record becomes a document of the collection whatever its
title (records are a list of (title, sequence): repeated titles do not overwrite each other) and the title is the
whole header line (fa_import.py:56); unlike it (first 500 records only, the file's last record dropped,
fa_import.py:22,43-55) every record of the file is kept.  normalise() also strips white space and upper-cases,
which the reference's importers leave to the GUI's input check (gui.py:62-69)."""
from __future__ import annotations

import xml.etree.ElementTree as ET

from .encoding import PackedSeqs, pack


def normalise(seq: str) -> str:
    return seq.strip().upper().replace("T", "U").replace("X", "N")


def read_seqxml(path: str) -> "list[tuple[str, str]]":
    """import_xml.py:4-17 — <entry id=...><RNAseq>...</RNAseq></entry>."""
    return [(entry.get("id"), normalise(entry.find("RNAseq").text)) for entry in ET.parse(path).getroot().findall("entry")]


def read_fasta(path: str) -> "list[tuple[str, str]]":
    """fa_import.py:39-62 — '>' header line (the whole line is the title), sequence possibly over several lines."""
    out: "list[tuple[str, str]]" = []
    name, parts = None, []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\r\n")
            if not line.strip():
                continue
            if line.startswith(">"):
                if name is not None:
                    out.append((name, normalise("".join(parts))))
                name, parts = line[1:], []
            else:
                parts.append(line.strip())
    if name is not None:
        out.append((name, normalise("".join(parts))))
    return out


class SequenceDB:
    """ids + packed sequences; `collection.find({})`-compatible so it can stand in for the Mongo
    collection the reference's callers pass around (IR:469)."""

    def __init__(self, records: "list[tuple[str, str]] | dict"):
        items = list(records.items()) if isinstance(records, dict) else list(records)
        self.ids = [k for k, _ in items]
        self.sequences = [v for _, v in items]
        self.packed: PackedSeqs = pack(self.sequences, bits=4)

    @classmethod
    def from_file(cls, path: str) -> "SequenceDB":
        return cls(read_seqxml(path) if path.lower().endswith(".xml") else read_fasta(path))

    def find(self, flt=None):
        return ({"_id": i, "sequence": s} for i, s in zip(self.ids, self.sequences))

    def __len__(self):
        return len(self.ids)
