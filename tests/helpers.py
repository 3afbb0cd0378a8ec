"""Shared test helpers (inputs from recipes, golden access)."""
from __future__ import annotations

import hashlib
import json
import os

from oracle import readgen
import recipes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def sha16(data: bytes) -> str:
    return hashlib.sha256(data).hexdigest()[:16]


def reads_for(recipe):
    """Inputs of a golden recipe, regenerated without the reference."""
    if recipe["kind"] == "explicit":
        return [tuple(r) for r in recipe["reads"]] if recipe["paired"] else list(recipe["reads"])
    genome = recipes.genome_text(recipe["genome"])
    return readgen.reference_style_reads(genome, recipe["L"], recipe["N"], recipe["paired"],
                                         d=recipe.get("d", 125), delta=recipe.get("delta", 0),
                                         seed=recipe["seed"])


def counts_sha(items) -> str:
    return sha16("".join("%s:%d\n" % kc for kc in sorted(items)).encode())
