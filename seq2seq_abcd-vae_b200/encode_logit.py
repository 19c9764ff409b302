# coding: utf-8
"""encode_logit.py of the reference (/root/reference ABCD-VAE/encode_logit.py:23-55): same CLI as encode.py, the CSV
carries the unnormalised category logits as columns `data_ix, dimension, logit`."""
from . import encode
from .encode import Encoder as _Encoder, get_parameters, rename_existing_file   # noqa: F401


class Encoder(_Encoder):
	def encode(self, data, is_packed=False, to_numpy=True, output='logit', return_argmax=False):
		return super(Encoder, self).encode(data, is_packed=is_packed, to_numpy=to_numpy, output=output, return_argmax=return_argmax)

	def encode_dataset(self, dataset, save_path, to_numpy=True, batch_size=1, output='logit', writer_threads=4):
		return super(Encoder, self).encode_dataset(dataset, save_path, to_numpy=to_numpy, batch_size=batch_size, output=output,
													writer_threads=writer_threads)


def main(argv=None):
	return encode.main(argv, output='logit')


if __name__ == '__main__':
	main()
