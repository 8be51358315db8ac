"""tests/golden/text_cleanup.json: the reference's own clean_text / score_sentence / select_best on a corpus of tricky strings
(run in the build container; needs /root/reference).   python oracle/pin_text_against_reference.py"""
import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, os.environ.get("VC_REFERENCE", "/root/reference"))
from core.postprocessing.text_cleaner import clean_text            # noqa: E402
from core.postprocessing.candidate_ranker import score_sentence, select_best   # noqa: E402

CORPUS = [
    "", "   ", "a man is playing a guitar", "a man is playing a guitar.", "Someone is sitting", "someone is sitting.", "a cat is sitting",
    "someone is sitting quietly", "a woman is is is cooking cooking in the kitchen", "------", "-- a dog runs", "====== .", "__ a boy jumps",
    "http://example.com watch this", "www.site.org a man", "<a href=x>link</a>", "Copyright 2020 by someone", '"quoted caption"', '"quoted caption".',
    "You are about to see a man", "Click here to subscribe", "subscribe to my channel", "Available on YouTube now", "watch live the game",
    "find out how a man cooks", "The video will show a man", "on the road again", "a man <b>bold</b> walks", "see reddit.com for more", "mailto:me@x.org",
    "a man is cooking click here", "a man report abuse", "dog video will be removed soon", "a man is walking in the U.S.A. today",
    "a man in the United States of America is running", "people in America are dancing in the street", "a man is standing in the front of a car",
    "a girl is in the middle of a field", "a boy at the side of the road", "a man is talking about how he cooks the food in the kitchen today",
    "a woman explains why the sky is blue", "A wonders of the world tour", "what is this", "that", "a man is riding a horse in 2019 near the big old barn by the river",
    "a man is riding a horse with ABC news logo near the big old barn by the river bank", "the U.S. army is marching down the long wide street in the city center now",
    "a man is driving a car model AB-12x down the long wide street in the city center", "a man is playing and then he is singing. a dog is barking loudly at the man in the yard.",
    "hello. a man is cooking food in the kitchen with a knife and a pan!  yes?", "a  man   is    slicing   a   tomato", "the the the man", "A MAN IS RUNNING",
    "a man is running 5 miles", "someone is sitting near a tree", "Someone is sitting on", "people are dancing", "is", "official facebook page of a man", "a b",
]

out = {"clean": [[s, clean_text(s)] for s in CORPUS], "score": [[s, score_sentence(s)] for s in CORPUS + [clean_text(s) for s in CORPUS]]}
cands = [[("S1", "a man is playing a guitar."), ("S2", "a man."), ("S3", "someone is sitting.")],
         [("S1", ""), ("S2", "Someone is in the scene."), ("S3", "a woman is slicing a tomato in the kitchen.")],
         [("S1", "a dog runs."), ("S2", "a dog runs."), ("S3", "a dog runs.")]]
out["select"] = [[c, list(select_best(c))] for c in cands]
(REPO / "tests" / "golden" / "text_cleanup.json").write_text(json.dumps(out, indent=0))
print(len(CORPUS), "strings pinned")
