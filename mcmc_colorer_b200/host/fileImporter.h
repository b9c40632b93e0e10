// Text edge-list reader, behaviour-compatible with the reference's utils/fileImporter.{h,cpp}:
// first line is a header and is skipped; every other non-empty line is "src dst weight" (whitespace separated);
// node names are mapped to ids in the iteration order of a std::unordered_set<std::string> (fileImporter.cpp:17,54-58).
#pragma once
#include <cstdint>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

class fileImporter {
public:
	fileImporter(std::string graphFileName, std::string labelFileName);
	~fileImporter();
	void fRewind();
	bool getNextEdge();

	std::ifstream graphFile;
	std::map<std::string, int> geneMap;
	std::map<int, std::string> inverseGeneMap;
	uint32_t nNodes{0};
	uint32_t nEdges{0};
	bool     edgeIsValid{false};
	uint32_t srcIdx{0}, dstIdx{0};
	double   edgeWgh{0};

private:
	std::stringstream ss;
	std::string inStr, src, dst;
	double ww_d{0};
};
