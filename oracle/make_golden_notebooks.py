#!/usr/bin/env python
"""Extracts the python-chess-ordered move lists the reference's own notebooks hold (build container only: needs
/root/reference) into tests/golden/notebook_traces.json.

* notebooks/verify_model.ipynb, cell `steps["steps"][:10]`: the first 10 plies of a self-play trace -- per ply the move
  played and the children of the root with their visit counts, in the order `predict` returned them, i.e. python-chess'
  legal-move generation order (src/chess.rs:665-676, src/mcts.rs:269-283) at TEN consecutive positions of a real game
  (20 - 36 legal moves, queen sorties, a pawn that can capture).
* notebooks/visualize_mcts.ipynb, cell `steps[:40]`: the 40 moves of another game (legality / replay check).
These are outputs of the reference running on real python-chess: the only such move-order pins that exist offline.
"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
NB = "/root/reference/notebooks"


def cell_output(nb_name, source_starts):
    nb = json.load(open(os.path.join(NB, nb_name)))
    for c in nb["cells"]:
        if "".join(c.get("source", [])).strip().startswith(source_starts):
            for o in c.get("outputs", []):
                txt = "".join(o.get("text", [])) if "text" in o else "".join(o.get("data", {}).get("text/plain", []))
                if txt:
                    return ast.literal_eval(txt)
    raise KeyError((nb_name, source_starts))


def main():
    steps = cell_output("verify_model.ipynb", 'steps["steps"][:10]')
    game = cell_output("visualize_mcts.ipynb", "steps[:40]")
    out = {
        "verify_model_trace": [{"move": s[0], "q": s[1], "children": [c[0] for c in s[2]], "visits": [c[1] for c in s[2]]}
                               for s in steps],
        "visualize_mcts_game": list(game),
        "source": 'notebooks/verify_model.ipynb (cell steps["steps"][:10]), notebooks/visualize_mcts.ipynb (cell steps[:40])',
    }
    dst = os.path.join(HERE, "..", "tests", "golden", "notebook_traces.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", dst, len(out["verify_model_trace"]), "plies with children,", len(out["visualize_mcts_game"]), "moves")


if __name__ == "__main__":
    main()
