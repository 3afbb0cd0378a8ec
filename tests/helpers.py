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


def splitmix_reads(recipe):
    """Reads of a "splitmix" recipe: oracle/readgen.py's numpy model of the device generator (ga_gen_genome /
    ga_gen_reads) -- uniform circular genome, every base substituted with probability sub_per_10k / 10000;
    pairs: mate 2 starts `dist` after mate 1."""
    genome = readgen.splitmix_genome_codes(recipe["G"], recipe["seed"])
    mates = recipe["N"] * (2 if recipe["paired"] else 1)
    codes = readgen.splitmix_reads_codes(genome, recipe["L"], 0, mates, recipe["seed"], recipe["sub_per_10k"],
                                         recipe["paired"], recipe.get("dist", 0))
    strings = readgen.codes_to_strings(codes)
    return list(zip(strings[0::2], strings[1::2])) if recipe["paired"] else strings


def reads_for(recipe):
    """Inputs of a golden recipe, regenerated without the reference."""
    if recipe["kind"] == "explicit":
        return [tuple(r) for r in recipe["reads"]] if recipe["paired"] else list(recipe["reads"])
    if recipe["kind"] == "splitmix":
        return splitmix_reads(recipe)
    genome = recipes.genome_text(recipe["genome"])
    return readgen.reference_style_reads(genome, recipe["L"], recipe["N"], recipe["paired"],
                                         d=recipe.get("d", 125), delta=recipe.get("delta", 0),
                                         seed=recipe["seed"])


def counts_sha(items) -> str:
    return sha16("".join("%s:%d\n" % kc for kc in sorted(items)).encode())
