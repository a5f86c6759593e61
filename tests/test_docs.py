"""The measurement files the documents cite exist (profiles/ is what the review reads)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cited(doc, prefix=""):
    text = open(os.path.join(ROOT, doc)).read()
    for m in re.finditer(r"`([\w\-./]+\.(?:json|txt|csv))`", text):
        path = m.group(1)
        if "/" not in path:
            path = prefix + path
        if path.startswith(("profiles/", "tests/golden/")):
            yield path


def test_cited_profile_files_exist():
    missing = []
    for doc, prefix in (("profiles/README.md", "profiles/"), ("DESIGN.md", ""), ("README.md", "")):
        for path in _cited(doc, prefix):
            if not (os.path.exists(os.path.join(ROOT, path)) or glob.glob(os.path.join(ROOT, path))):
                missing.append((doc, path))
    assert missing == [], missing
