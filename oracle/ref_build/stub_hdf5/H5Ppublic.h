/* empty stub: HDF5 is absent in this image; only cooling table readers (disabled) use it */
