"""The LLM prompt of a navigation step (Policy_Dynam3D_VLN.py:436): pure string formatting, no dependencies.

`<image>` is repeated once per spliced visual token (576 patch tokens of the view the LLM sees + instance + zone tokens); POL:456 then
replaces the embeddings of exactly those positions (after the first two ids) by the projected 3D tokens."""

IMAGE_TOKEN = "<image>"


def build_prompt(n_image_tokens, instruction, history):
    return ("<|user|>\n" + IMAGE_TOKEN * n_image_tokens + "\nInstruction:\n" + instruction + "\nHistory actions:\n" + "".join(history) +
            "<|end|>\n<|assistant|>\nNext action:\n")
