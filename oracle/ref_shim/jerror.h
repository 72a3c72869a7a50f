/* stub, see jpeglib.h */
