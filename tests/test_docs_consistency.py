"""Docs must not point at files that do not exist (profiles, tests, sources)."""
from __future__ import annotations

import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
DOCS = ["DESIGN.md", "INTEGRATION.md", "README.md", "profiles/README.md", "benchmarks/README.md"]
PREFIXES = ("profiles/", "tests/", "oracle/", "benchmarks/", "include/", "k8s_cc_manager_b200/", "deployments/")


@pytest.mark.parametrize("doc", DOCS)
def test_referenced_paths_exist(doc):
    text = (ROOT / doc).read_text()
    missing = []
    for token in set(re.findall(r"`([^`\s]+)`", text)):
        token = token.split("::")[0].rstrip(".,;:)")
        candidates = []
        if token.startswith(PREFIXES):
            candidates.append(token)
        elif doc.startswith("profiles/") and re.fullmatch(r"r[12][a-z]?_[\w.{},-]+\.(json|csv|log|txt|md)", token):
            candidates.append("profiles/" + token)
        elif doc.startswith("benchmarks/") and re.fullmatch(r"[\w-]+\.(py|sh|cu)", token):
            candidates.append("benchmarks/" + token)
        for c in candidates:
            if "{" in c or "*" in c or "<" in c or "…" in c:
                continue
            if c == "oracle/_ref":          # named only to say it does not apply to this reference
                continue
            if not (ROOT / c).exists() and not (ROOT / Path(c).name).exists():
                missing.append(c)
    assert not missing, f"{doc} references missing files: {sorted(missing)}"
