// Stand-in for the CMake-generated export header: static build, default visibility.
#ifndef LOGUTIL_EXPORT_H
#define LOGUTIL_EXPORT_H
#define logutil_API
#define LOGUTIL_NO_EXPORT
#endif
