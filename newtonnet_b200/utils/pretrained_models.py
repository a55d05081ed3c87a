"""Mirror of newtonnet/utils/pretrained_models.py:7-43 (signature only: this build has no network)."""
import os

__all__ = ['download_checkpoint']


def download_checkpoint(model: str) -> str:
    cache = os.path.join(os.path.expanduser('~'), '.cache', 'newtonnet', f'{model}.pt')
    if os.path.exists(cache):
        return cache
    raise RuntimeError(f"pretrained checkpoint '{model}' is not cached at {cache} and downloading is disabled "
                       f"in this build; pass a local checkpoint path instead")
