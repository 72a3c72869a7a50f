/* TEST INFRASTRUCTURE ONLY.  The reference includes <jerror.h> next to <jpeglib.h> but uses none of
 * its message codes (only JMSG_LENGTH_MAX, which jpeglib.h defines); nothing to declare. */
#ifndef UHDR_ORACLE_JERROR_STUB_H
#define UHDR_ORACLE_JERROR_STUB_H
#endif
