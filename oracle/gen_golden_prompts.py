"""ORACLE tooling — extracts the prompt templates of the reference trainers (string constants, read with `ast`; nothing is
imported or executed) into tests/golden/prompt_templates.json, so the drop-in's prompts can be checked character by
character on a box without the reference.  Build container only:  python oracle/gen_golden_prompts.py"""
import ast
import json
import os

REF = "/root/reference/src/time_r1/rl"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "prompt_templates.json")


def constants(path):
    out = {}
    for node in ast.parse(open(path).read()).body:
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and \
                node.targets[0].id.startswith(("QUESTION_TEMPLATE", "SYSTEM_PROMPT")):
            out[node.targets[0].id] = ast.literal_eval(node.value)
    return out


if __name__ == "__main__":
    data = {"timer1_trainer.py": constants(os.path.join(REF, "timer1_trainer.py")),
            "timer1_trainer_ft.py": constants(os.path.join(REF, "timer1_trainer_ft.py"))}
    json.dump(data, open(OUT, "w"), indent=0, sort_keys=True)
    print({k: sorted(v) for k, v in data.items()})
