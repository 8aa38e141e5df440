#!/usr/bin/env python3
"""Copies the raw texts of the reference's sample corpus (doc/samples/texts/HSE rules/*.txt, 30 short Russian documents,
DATA not code) into tests/golden/hse_texts.json.  Their strings collections as the REFERENCE's own
utils.text_to_strings_collection produced them are already in golden.json (hse.docs[*].strings, written by make_golden.py),
so the preprocessing -- on the host and on the device -- can be checked against the reference from raw text.
usage (in the authoring container, where /root/reference exists): python tests/golden/make_hse_texts.py"""
import json
import os

REF = "/root/reference/doc/samples/texts/HSE rules"
HERE = os.path.dirname(os.path.abspath(__file__))

texts = {}
for fn in sorted(os.listdir(REF)):
    if fn.endswith(".txt"):
        with open(os.path.join(REF, fn), "rb") as f:
            texts[fn] = f.read().decode("utf-8")
with open(os.path.join(HERE, "hse_texts.json"), "w", encoding="utf-8") as f:
    json.dump(texts, f, ensure_ascii=False, indent=0, sort_keys=True)
print("%d texts, %d characters" % (len(texts), sum(len(t) for t in texts.values())))
