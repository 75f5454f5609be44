"""dump / load (reference: surprise/dump.py:8-58): pickle of {'predictions', 'algo'}.  Fitted algorithms
drop their device handles in __getstate__ and keep numpy float64 attributes, so they pickle."""
import pickle


def dump(file_name, predictions=None, algo=None, verbose=0):
    with open(file_name, "wb") as fh:
        pickle.dump({"predictions": predictions, "algo": algo}, fh, protocol=pickle.HIGHEST_PROTOCOL)
    if verbose:
        print("The dump has been saved as file", file_name)


def load(file_name):
    with open(file_name, "rb") as fh:
        d = pickle.load(fh)
    return d["predictions"], d["algo"]
