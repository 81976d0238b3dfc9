// Stand-in for the CMake-generated export header: static build, default visibility.
#ifndef OHMUTIL_EXPORT_H
#define OHMUTIL_EXPORT_H
#define ohmutil_API
#define OHMUTIL_NO_EXPORT
#endif
